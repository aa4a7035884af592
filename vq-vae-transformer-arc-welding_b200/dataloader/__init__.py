from .latentspace_dataloader import (CycleIdCache, LatentSpaceEncoder, OnTheFlyTokenizer, bulk_encode_ids,  # noqa: F401
                                     gather_sharded, latent_dataset_name, reduce_counts, save_latent_dataset, shard_range)
