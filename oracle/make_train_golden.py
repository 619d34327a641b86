"""TEST INFRASTRUCTURE ONLY.  tests/golden/mini_train.pt: the UNMODIFIED reference (oracle/ref_loader.py) run with
``return_loss=True`` (reference M:1722-1896, use_gan=False, perceptual_loss_weight=0) on the `mini` config:

* eval mode:  total_loss, recon_loss (aux terms are zero in eval, A.1 step 6) and ``return_recon_loss_only``;
* train mode: total_loss, LossBreakdown (recon, aux, (per_sample_entropy, batch_entropy, commitment)) and, after
  ``total_loss.backward()``, the gradient of EVERY parameter (what the trainer's step consumes, T:356-363).

Runs only in the build container:   python -m oracle.make_train_golden
"""
from __future__ import annotations

import os

import torch

from oracle import weights as W
from oracle.make_golden import CONFIGS, GOLDEN_DIR
from oracle.ref_loader import build_reference_tokenizer


def grad_digest(g: torch.Tensor, max_elems: int = 1024):
    flat = g.reshape(-1)
    stride = max(1, -(-flat.numel() // max_elems))
    return dict(norm=float(flat.double().norm()), stride=stride, sample=flat[::stride].clone(), shape=tuple(g.shape))


def make(name="mini_train", base="mini", vseed=1234):
    cfg = CONFIGS[base]
    kwargs = dict(cfg["kwargs"], use_gan=False, perceptual_loss_weight=0.)
    torch.manual_seed(0)
    model = build_reference_tokenizer(**kwargs)
    W.fill_state_dict_(model, cfg["wseed"])
    video = W.synth_video(*cfg["video"][:3], cfg["video"][3], seed=vseed)
    out = dict(name=name, kwargs=kwargs, video_shape=tuple(cfg["video"]), wseed=cfg["wseed"], vseed=vseed)
    ckw = {}
    if kwargs.get("dim_cond") is not None:          # conditioned layers (cond_residual): a (B, dim_cond) vector per clip
        gen = torch.Generator(device="cpu")
        gen.manual_seed(cfg["cseed"])
        out["cond"] = torch.randn(cfg["video"][0], kwargs["dim_cond"], generator=gen)
        ckw = dict(cond=out["cond"])

    model.eval()
    with torch.no_grad():
        total, bd = model(video, return_loss=True, **ckw)
        rl, recon = model(video, return_recon_loss_only=True, **ckw)
    out["eval"] = dict(total_loss=total.clone(), recon_loss=bd.recon_loss.clone(), aux=torch.as_tensor(bd.lfq_aux_loss).clone(),
                       recon_loss_only=rl.clone(), recon_mean=recon.mean(dim=(3, 4)).clone())

    model.train()
    for p in model.parameters():
        p.grad = None
    total, bd = model(video, return_loss=True, **ckw)
    total.backward()
    # per parameter: the gradient's L2 norm and a strided sample of at most ~1024 elements (the full set is 11 MB)
    grads = {k: (grad_digest(p.grad.detach()) if p.grad is not None else None) for k, p in model.named_parameters()}
    qlb = bd.quantizer_loss_breakdown
    out["train"] = dict(total_loss=total.detach().clone(), recon_loss=bd.recon_loss.detach().clone(),
                        aux=torch.as_tensor(bd.lfq_aux_loss).detach().clone(), grads=grads)
    if qlb is not None:           # LFQ only (FSQ has no auxiliary loss, M:1700-1703)
        out["train"].update(per_sample_entropy=qlb.per_sample_entropy.detach().clone(),
                            batch_entropy=qlb.batch_entropy.detach().clone(), commitment=qlb.commitment.detach().clone())
    out["reference_commit"] = "a00519fa (v0.5.1)"
    out["third_party"] = "oracle/shims (restated LFQ/TaylorSeriesLinearAttn; real packages unavailable)"
    path = os.path.join(GOLDEN_DIR, f"{name}.pt")
    torch.save(out, path)
    n_none = sum(g is None for g in grads.values())
    gn = sum(g["norm"] ** 2 for g in grads.values() if g is not None) ** 0.5
    print(f"[golden] {name}: eval total {out['eval']['total_loss'].item():.6f}; train total {out['train']['total_loss'].item():.6f} "
          f"recon {out['train']['recon_loss'].item():.6f} aux {float(out['train']['aux']):.6f}; {len(grads)} params "
          f"({n_none} without grad), |grad| = {gn:.4e}; {os.path.getsize(path) / 1e3:.0f} KB")


if __name__ == "__main__":
    make()
    make("mini_mc_train", base="mini_mc", vseed=1238)       # num_codebooks = 2, lfq_spherical
    make("mini_fsq_train", base="mini_fsq", vseed=1234)     # FSQ: straight-through round, no auxiliary loss
    make("mini_gateloop_train", base="mini_gateloop", vseed=1237)
    make("mini_cond_train", base="mini_cond", vseed=1249)   # cond_residual (ResidualUnitMod / Conv3DMod) + the cond stems
    make("mini_sff_train", base="mini_sff", vseed=1234)     # separate_first_frame_encoding
    make("pad_reflect_train", base="pad_reflect", vseed=1236)       # pad_mode of conv_in / conv_out (M:925-927)
    make("pad_circular_train", base="pad_circular", vseed=1236)
