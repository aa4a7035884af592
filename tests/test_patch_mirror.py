"""The VQ-VAE-Patch mirror against fixtures produced by the unmodified reference model
(oracle/make_golden.py: model/vq_vae_patch_embedd.py).  CPU part: state-dict layout and the
encoder restatement (patchify + per-token residual MLP + projection)."""
import numpy as np
import pytest
import torch

import cases as C
import vqb200


def _build(case):
    return vqb200.VQVAEPatch(hidden_dim=case["hidden_dim"], input_dim=case["input_dim"],
                             num_embeddings=case["num_embeddings"], embedding_dim=case["embedding_dim"],
                             n_resblocks=case["n_resblocks"], learning_rate=1e-3, dropout_p=0.0,
                             patch_size=case["patch_size"], seq_len=case["seq_len"],
                             batch_norm=case["batch_norm"], beta=case["beta"])


def _state_dict(patch_golden, name):
    pre = f"{name}/sd/"
    return {k[len(pre):]: torch.from_numpy(patch_golden[k]) for k in patch_golden.files if k.startswith(pre)}


@pytest.mark.parametrize("case", C.PATCH_CASES, ids=[c["name"] for c in C.PATCH_CASES])
def test_reference_state_dict_loads_strictly(case, patch_golden, manifest):
    model = _build(case)
    sd = _state_dict(patch_golden, case["name"])
    assert list(model.state_dict().keys()) == manifest["patch"][case["name"]]["keys"]
    model.load_state_dict(sd, strict=True)
    for k, v in model.state_dict().items():
        assert tuple(v.shape) == tuple(sd[k].shape)


@pytest.mark.parametrize("bn", [False, True])
def test_default_config_key_layout(bn, manifest):
    want = manifest["default_config_state_dict"][f"batch_norm={bn}"]
    model = vqb200.VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32,
                              n_resblocks=8, learning_rate=1e-3, dropout_p=0.1, patch_size=25, batch_norm=bn)
    got = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert got == want
    assert sum(p.numel() for p in model.parameters()) == manifest["default_config_state_dict"][f"n_params_batch_norm={bn}"]
    if not bn:
        assert len(got) == 80 and "vector_quantization.embedding.weight" in got


@pytest.mark.parametrize("case", C.PATCH_CASES, ids=[c["name"] for c in C.PATCH_CASES])
def test_encoder_restatement_matches_reference(case, patch_golden, manifest):
    name = case["name"]
    model = _build(case)
    model.load_state_dict(_state_dict(patch_golden, name), strict=True)
    model.eval()
    x = torch.from_numpy(C.make_cycles(case))
    with torch.no_grad():
        tokens = model.patch_embed(x)
        z_e = model.encoder(tokens)
    np.testing.assert_allclose(tokens.numpy(), patch_golden[f"{name}/tokens"], rtol=1e-5, atol=1e-6)
    assert list(z_e.shape) == manifest["patch"][name]["z_e_shape"]
    # the reference returns these values as a permuted view (manifest z_e_strides); the mirror writes
    # contiguous (B, T, D) rows, the layout the tcgen05 quantiser streams without a packing copy
    assert z_e.is_contiguous() and manifest["patch"][name]["z_e_strides"] != list(z_e.stride())
    np.testing.assert_allclose(z_e.contiguous().numpy(), patch_golden[f"{name}/z_e"], rtol=2e-5, atol=2e-6)


def test_training_mode_batchnorm_uses_per_position_statistics(patch_golden):
    """With BatchNorm in training mode the reference normalises every position with its own
    batch statistics (model/vq_vae_patch_embedd.py:106-110); the mirror must not fuse them."""
    case = C.PATCH_CASES[1]
    assert case["batch_norm"]
    model = _build(case)
    model.load_state_dict(_state_dict(patch_golden, case["name"]), strict=True)
    model.train()
    x = torch.from_numpy(C.make_cycles(case))
    tokens = model.patch_embed(x)
    out = model.encoder[0](tokens)
    bn = model.encoder[0].shared_conv[0].block[2]
    assert int(bn.num_batches_tracked) == tokens.shape[2]      # one update per position
    pos0 = model.encoder[0].shared_conv(tokens[:, :, 0:1].detach())
    assert out.shape == tokens.shape and pos0.shape[2] == 1
