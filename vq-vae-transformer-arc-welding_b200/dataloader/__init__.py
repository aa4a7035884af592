from .latentspace_dataloader import (LatentSpaceEncoder, OnTheFlyTokenizer, bulk_encode_ids, gather_sharded, reduce_counts,  # noqa: F401
                                     shard_range)
