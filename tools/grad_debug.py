"""Developer tool: per-parameter gradient cosine (CUDA path vs CPU oracle with bf16 emulation), in module order,
for training-mode and eval-mode BatchNorm."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from box2mask_b200.model import Model  # noqa: E402
from box2mask_b200.selection_net import default_config  # noqa: E402
from box2mask_b200.synthetic import label_maps, make_batch  # noqa: E402
from oracle.selection_net import OracleNet, detection_loss, seeded_state_dict  # noqa: E402


def cos(a, b):
    return float(torch.nn.functional.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))


cfg = default_config(mlp_bb_scores_start_epoch=0)
valid, id2idx, is_fg = label_maps(20)
model = Model(cfg, valid, id2idx, None, is_fg, device="cuda")
sd = seeded_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()})
batch = make_batch(8, seed=5, scale=0.2, density=1.2e4)
for training in (False, True):
    model.load_state_dict(sd)
    model.train() if training else model.eval()
    for p in model.parameters():
        p.grad = None
    losses, pred = model.compute_loss_detection(batch, epoch=0)
    losses["optimization_loss"].backward()
    osd = {k: v.clone() for k, v in sd.items()}
    for k, v in osd.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    out = OracleNet(osd, cfg, training=training, emulate_bf16=True).forward(
        batch["vox_coords"].numpy(), batch["vox_features"], batch["pooling_ids"])
    ol = detection_loss(out, batch, cfg, 0, id2idx)
    ol["optimization_loss"].backward()
    print("==== training=%s loss ours %.5f oracle %.5f" % (training, float(losses["optimization_loss"]), float(ol["optimization_loss"])))
    for head in cfg.network_heads:
        print("  head %-16s cos %.5f" % (head, cos(pred[head].detach().cpu(), out[head].detach())))
    for name, p in model.net.named_parameters():
        if p.grad is None:
            print("  %-44s NO GRAD" % name)
            continue
        g, r = p.grad.cpu(), osd[name].grad
        print("  %-44s cos %.4f  |g| %.3e |ref| %.3e" % (name, cos(g, r), float(g.norm()), float(r.norm())))
