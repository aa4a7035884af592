"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE.

TEST INFRASTRUCTURE.  Runs only in the authoring container, where
/root/reference exists; the GPU box and the test-suite only read the committed
fixtures.  `lightning` and `vector_quantize_pytorch` are not installed, so the
two import-time dependencies of model/vector_quantizer.py:4,6 are satisfied by
the stand-ins under oracle/ref_stub/ (SURVEY.md section 8(c)); the reference
files themselves are imported as they are, from where they lie.

    python oracle/make_golden.py            # rewrites tests/golden/vq_golden.npz etc.

Settings that matter: CPU, fp32,
torch.set_float32_matmul_precision('highest'), eval() for forward fixtures.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("VQ_REFERENCE_ROOT", "/root/reference")
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, GOLDEN)

import cases as C  # noqa: E402


def load_reference():
    if not os.path.isdir(REF):
        raise FileNotFoundError(f"{REF} not present: golden vectors can only be regenerated where the reference is mounted")
    sys.path[:0] = [os.path.join(HERE, "ref_stub"), REF]
    from model.vector_quantizer import VectorQuantizer  # type: ignore
    from model.vq_vae_patch_embedd import VQVAEPatch  # type: ignore
    return VectorQuantizer, VQVAEPatch


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def run_vq_cases(VectorQuantizer):
    out = {}
    meta = {}
    for case in C.CASES:
        name = case["name"]
        storage, logical, E = C.make_inputs(case)
        vq = VectorQuantizer(case["K"], case["D"], case["beta"])
        with torch.no_grad():
            vq.embedding.weight.copy_(torch.from_numpy(E))
        st = torch.from_numpy(storage)
        z = st.permute(0, 2, 1) if case["layout"] == "permuted" else st
        assert tuple(z.shape) == tuple(case["shape"]) and np.array_equal(z.numpy(), logical, equal_nan=True)
        z = z.detach().requires_grad_(True)
        loss, zq, ppl, onehot, idx = vq(z)
        n = idx.shape[0]
        assert onehot.shape == (n, case["K"]) and idx.shape == (n, 1) and idx.dtype == torch.int64
        chk = torch.zeros(n, case["K"]).scatter_(1, idx, 1)
        assert torch.equal(chk, onehot)
        assert zq.is_contiguous() and zq.shape == z.shape
        zq_np = zq.detach().numpy().reshape(n, case["D"])
        out[f"{name}/idx"] = idx.numpy().reshape(-1).astype(np.int32)
        out[f"{name}/loss"] = np.float32(loss.item())
        out[f"{name}/perplexity"] = np.float32(ppl.item())
        out[f"{name}/zq_head"] = zq_np[: C.GRAD_ROWS].copy()
        m = dict(n=int(n), zq_sha256=sha(zq_np), zq_requires_grad=bool(zq.requires_grad),
                 loss_requires_grad=bool(loss.requires_grad), ppl_requires_grad=bool(ppl.requires_grad))
        if case["bwd"]:
            w = torch.from_numpy(C.upstream_weights(case))
            total = C.G_LOSS * loss + (w * zq).sum()
            total.backward()
            gz = z.grad.detach().numpy().reshape(n, case["D"])
            out[f"{name}/grad_z_head"] = gz[: C.GRAD_ROWS].copy()
            out[f"{name}/grad_z_colsum"] = gz.astype(np.float64).sum(0)
            out[f"{name}/grad_E"] = vq.embedding.weight.grad.detach().numpy().copy()
        meta[name] = m
        print(f"{name:18s} n={n:6d} loss={loss.item():.6e} ppl={ppl.item():.4f}")
    return out, meta


def run_patch_cases(VQVAEPatch):
    out = {}
    meta = {}
    for case in C.PATCH_CASES:
        name = case["name"]
        torch.manual_seed(case["seed"])
        model = VQVAEPatch(hidden_dim=case["hidden_dim"], input_dim=case["input_dim"],
                           num_embeddings=case["num_embeddings"], embedding_dim=case["embedding_dim"],
                           n_resblocks=case["n_resblocks"], learning_rate=1e-3, dropout_p=0.0,
                           patch_size=case["patch_size"], seq_len=case["seq_len"],
                           batch_norm=case["batch_norm"], beta=case["beta"])
        with torch.no_grad():  # spread the codebook so that several codes are used
            model.vector_quantization.embedding.weight.mul_(case["num_embeddings"] * 0.5)
        x = torch.from_numpy(C.make_cycles(case))
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        model.eval()
        with torch.no_grad():
            tokens = model.patch_embed(x)
            z_e = model.encoder(tokens)
            loss, zq, ppl, _, idx = model.vector_quantization(z_e)
            emb_loss, x_hat, ppl2 = model(x)
        assert torch.equal(loss, emb_loss)
        for k, v in sd.items():
            out[f"{name}/sd/{k}"] = v.numpy()
        out[f"{name}/tokens"] = tokens.numpy()
        out[f"{name}/z_e"] = z_e.contiguous().numpy()
        out[f"{name}/idx"] = idx.numpy().reshape(-1).astype(np.int32)
        out[f"{name}/zq"] = zq.numpy()
        out[f"{name}/emb_loss"] = np.float32(emb_loss.item())
        out[f"{name}/perplexity"] = np.float32(ppl.item())
        out[f"{name}/x_hat"] = x_hat.numpy()
        # one training-mode step (dropout_p = 0 so it is deterministic): total loss and two gradients
        model.train()
        model.zero_grad()
        total, recon, _ = model._forward_setp(x)
        total.backward()
        out[f"{name}/train_total"] = np.float32(total.item())
        out[f"{name}/train_recon"] = np.float32(recon.item())
        out[f"{name}/grad_codebook"] = model.vector_quantization.embedding.weight.grad.numpy().copy()
        out[f"{name}/grad_enc_proj"] = model.encoder[1].shared_conv.weight.grad.numpy().copy()
        out[f"{name}/grad_patch_proj"] = model.patch_embed.proj.weight.grad.numpy().copy()
        meta[name] = dict(z_e_strides=list(z_e.stride()), z_e_shape=list(z_e.shape),
                          enc_out_len=int(model.enc_out_len), keys=list(sd.keys()))
        print(f"{name:18s} emb_loss={emb_loss.item():.6e} used_codes={len(set(idx.reshape(-1).tolist()))}")
    return out, meta


def build_ref_model(VQVAEPatch, case):
    torch.manual_seed(case["seed"])
    model = VQVAEPatch(hidden_dim=case["hidden_dim"], input_dim=case["input_dim"],
                       num_embeddings=case["num_embeddings"], embedding_dim=case["embedding_dim"],
                       n_resblocks=case["n_resblocks"], learning_rate=1e-3, dropout_p=0.0,
                       patch_size=case["patch_size"], seq_len=case["seq_len"],
                       batch_norm=case["batch_norm"], beta=case["beta"])
    with torch.no_grad():  # spread the codebook so that several codes are used
        model.vector_quantization.embedding.weight.mul_(case["num_embeddings"] * 0.5)
    return model


def run_wide_case(VQVAEPatch):
    """A model the fused tcgen05 encoder layers accept (hidden_dim = 256): the reference's z_e and ids, stored with
    only the tensors its encode path reads (tests/golden/cases.py: PATCH_WIDE_CASE)."""
    case = C.PATCH_WIDE_CASE
    name = case["name"]
    model = build_ref_model(VQVAEPatch, case).eval()
    x = torch.from_numpy(C.make_cycles(case))
    with torch.no_grad():
        tokens = model.patch_embed(x)
        z_e = model.encoder(tokens)
        _loss, _zq, ppl, _, idx = model.vector_quantization(z_e)
    sd = model.state_dict()
    out = {}
    for k, v in sd.items():
        if k.startswith("patch_embed.") or k.startswith("encoder.1.") or k.startswith("vector_quantization."):
            out[f"{name}/sd/{k}"] = v.numpy().copy()
        elif k.startswith("encoder.0.") and k.endswith(".bias"):
            out[f"{name}/sd/{k}"] = v.numpy().copy()
        elif k.startswith("encoder.0.") and k.endswith(".weight"):
            assert v.shape[2] == 3
            out[f"{name}/centre/{k}"] = v[:, :, 1].numpy().copy()
    out[f"{name}/z_e"] = z_e.contiguous().numpy()
    out[f"{name}/idx"] = idx.numpy().reshape(-1).astype(np.int32)
    out[f"{name}/perplexity"] = np.float32(ppl.item())
    meta = dict(z_e_shape=list(z_e.shape), used_codes=len(set(idx.reshape(-1).tolist())))
    print(f"{name:18s} tokens={idx.numel()} used_codes={meta['used_codes']} ppl={ppl.item():.3f}")
    return out, meta


def run_bulk_case(VQVAEPatch):
    """The reference's own bulk loops (dataloader/latentspace_dataloader.py:171-263) run as unbound functions on a
    stand-in `self` (the constructor wants the ASIMoW CSV data set): the three arrays they build from overlapping
    windows."""
    import types
    from dataloader.latentspace_dataloader import LatentSpaceDataLoader as L  # type: ignore
    case = C.BULK_CASE
    name = case["name"]
    mcase = next(c for c in C.PATCH_CASES if c["name"] == case["model"])
    model = build_ref_model(VQVAEPatch, mcase)
    win, labels, slices = C.make_windows(case)
    win_t, lab_t = torch.from_numpy(win), torch.from_numpy(labels)
    loader = [(win_t[lo:hi], lab_t[lo:hi]) for lo, hi in slices]
    me = types.SimpleNamespace(latent_space_model=model, device="cpu", window_size=200, task="classification_ids")
    me.get_latent_space = types.MethodType(L.get_latent_space, me)
    me.get_latent_space_IDs = types.MethodType(L.get_latent_space_IDs, me)
    me.create_latent_space_dataset_VQ_VAE_IDs = types.MethodType(L.create_latent_space_dataset_VQ_VAE_IDs, me)
    ids, y = L.create_latent_space_dataset_VQ_VAE_IDs(me, loader, seq_len=case["seq_len"], has_patch_embed=True)
    zq, y2 = L.create_latent_space_dataset_VQ_VAE(me, loader, seq_len=case["seq_len"], has_patch_embed=True)
    me.task = "autoregressive_ids"
    ar, y3 = L.create_latent_space_dataset_VQ_VAE_autoreggressive(me, [w for w, _ in loader], seq_len=case["seq_len"],
                                                                 has_patch_embed=True)
    me.task = "autoregressive_ids_classification"
    ar2, y4 = L.create_latent_space_dataset_VQ_VAE_autoreggressive(me, loader, seq_len=case["seq_len"],
                                                                  has_patch_embed=True)
    assert np.array_equal(ar2, ar) and np.array_equal(y4, y)
    # the transformer's data set built from those ids by the reference's own class (dataloader/base_dataloader.py:74-110)
    from dataloader.base_dataloader import MyLatentAutoregressiveDataset  # type: ignore
    ds = MyLatentAutoregressiveDataset(ar, y)
    items = [ds[i] for i in range(len(ds))]
    ar_x = np.stack([it[0].numpy() for it in items])
    ar_cond = np.stack([it[1].numpy() for it in items])
    ar_y = np.stack([it[2].numpy() for it in items])
    # the encoder outputs behind those ids (same reference modules, same per-cycle slices): what a near-tie is judged on
    model.eval()
    with torch.no_grad():
        z_e = np.stack([np.stack([model.encoder(model.patch_embed(w[:, i * 200:(i + 1) * 200, :].clone())).contiguous().numpy()
                                  for i in range(case["seq_len"])], axis=1) for w, _ in loader])
    z_e = z_e.reshape((-1,) + z_e.shape[2:]) if z_e.ndim == 6 else np.concatenate(list(z_e), axis=0)
    out = {f"{name}/ids": ids, f"{name}/labels": y, f"{name}/zq": zq, f"{name}/labels_zq": y2,
           f"{name}/ar_ids": ar, f"{name}/ar_labels": y3, f"{name}/z_e": z_e.astype(np.float32),
           f"{name}/ar_ds_x": ar_x, f"{name}/ar_ds_cond": ar_cond, f"{name}/ar_ds_y": ar_y,
           f"{name}/ar_ds_num_classes": np.int64(ds.num_classes)}
    meta = {k.split("/")[1]: dict(shape=list(v.shape), dtype=str(v.dtype)) for k, v in out.items()}
    print(f"{name:18s} ids {ids.shape} {ids.dtype}  zq {zq.shape} {zq.dtype}  ar {ar.shape}")
    return out, meta


def run_sequence_case(VQVAEPatch):
    """The reference's window builder (`ASIMoWDataLoader.create_sequence_ds`, dataloader/asimow_dataloader.py:185-206: windows
    of seq_len consecutive cycles with a stride of one cycle, n - seq_len of them, label of window i = y[i + seq_len]) on a
    stream of cycles, its data-set class's float32 cast (dataloader/base_dataloader.py:29), and the reference's id loop
    (dataloader/latentspace_dataloader.py:205-243) over those windows -- what a builder that encodes every cycle ONCE and
    windows the ids has to reproduce."""
    import types
    from dataloader.asimow_dataloader import ASIMoWDataLoader as A  # type: ignore
    from dataloader.latentspace_dataloader import LatentSpaceDataLoader as L  # type: ignore
    case = C.BULK_CASE
    name = case["name"]
    mcase = next(c for c in C.PATCH_CASES if c["name"] == case["model"])
    model = build_ref_model(VQVAEPatch, mcase)
    stream, cycle_labels = C.make_stream(case)
    wx, wy = A.create_sequence_ds(types.SimpleNamespace(window_size=200, window_offset=0), stream.astype(np.float64),
                                  cycle_labels, case["seq_len"])
    win_t, lab_t = torch.tensor(wx, dtype=torch.float32), torch.tensor(wy, dtype=torch.long)
    loader = [(win_t[lo:lo + case["batch"]], lab_t[lo:lo + case["batch"]]) for lo in range(0, len(win_t), case["batch"])]
    me = types.SimpleNamespace(latent_space_model=model, device="cpu", window_size=200, task="classification_ids")
    me.get_latent_space = types.MethodType(L.get_latent_space, me)
    me.get_latent_space_IDs = types.MethodType(L.get_latent_space_IDs, me)
    ids, y = L.create_latent_space_dataset_VQ_VAE_IDs(me, loader, seq_len=case["seq_len"], has_patch_embed=True)
    zq, y2 = L.create_latent_space_dataset_VQ_VAE(me, loader, seq_len=case["seq_len"], has_patch_embed=True)
    assert np.array_equal(y, y2)
    out = {f"{name}/seq_ids": ids, f"{name}/seq_labels": y, f"{name}/seq_zq": zq}
    meta = {k.split("/")[1]: dict(shape=list(v.shape), dtype=str(v.dtype)) for k, v in out.items()}
    print(f"{name:18s} create_sequence_ds windows {wx.shape} -> ids {ids.shape} labels {y.shape}")
    return out, meta


def default_config_keys(VQVAEPatch):
    """State-dict layout of the repo-default model (train_reconstruction_embedding.py:220-230)."""
    res = {}
    for bn in (False, True):
        torch.manual_seed(0)
        model = VQVAEPatch(hidden_dim=512, input_dim=2, num_embeddings=256, embedding_dim=32,
                           n_resblocks=8, learning_rate=1e-3, dropout_p=0.1, patch_size=25,
                           batch_norm=bn)
        sd = model.state_dict()
        res[f"batch_norm={bn}"] = {k: list(v.shape) for k, v in sd.items()}
        res[f"n_params_batch_norm={bn}"] = int(sum(p.numel() for p in model.parameters()))
    return res


def main():
    torch.set_float32_matmul_precision("highest")
    torch.manual_seed(0)
    VectorQuantizer, VQVAEPatch = load_reference()
    vq_out, vq_meta = run_vq_cases(VectorQuantizer)
    np.savez_compressed(os.path.join(GOLDEN, "vq_golden.npz"), **vq_out)
    p_out, p_meta = run_patch_cases(VQVAEPatch)
    np.savez_compressed(os.path.join(GOLDEN, "patch_golden.npz"), **p_out)
    w_out, w_meta = run_wide_case(VQVAEPatch)
    np.savez_compressed(os.path.join(GOLDEN, "patch_wide_golden.npz"), **w_out)
    b_out, b_meta = run_bulk_case(VQVAEPatch)
    s_out, s_meta = run_sequence_case(VQVAEPatch)
    b_out.update(s_out)
    b_meta.update(s_meta)
    np.savez_compressed(os.path.join(GOLDEN, "bulk_golden.npz"), **b_out)
    manifest = dict(
        generator="oracle/make_golden.py",
        reference_root=REF,
        torch=torch.__version__,
        numpy=np.__version__,
        threads=torch.get_num_threads(),
        matmul_precision=torch.get_float32_matmul_precision(),
        vq=vq_meta,
        patch=p_meta,
        patch_wide=w_meta,
        bulk=b_meta,
        default_config_state_dict=default_config_keys(VQVAEPatch),
    )
    with open(os.path.join(GOLDEN, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    for fn in ("vq_golden.npz", "patch_golden.npz", "patch_wide_golden.npz", "bulk_golden.npz", "manifest.json"):
        print(fn, os.path.getsize(os.path.join(GOLDEN, fn)), "bytes")


if __name__ == "__main__":
    main()
