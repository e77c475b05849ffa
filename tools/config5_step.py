"""BASELINE.json config 5 on one GPU: a ContrastivePretraining-style training step with
HuggingFace CLIP towers (ViT-B/16 vision + CLIP text transformer, random init, synthetic
image/text batch, bf16 autocast), once with the reference's loss arithmetic in plain torch
(F.normalize -> logit_scale * a @ b.T -> cross_entropy x2, contrastive.py:89-144) and once with
the fused loss behind ``mmlearn_b200.task.ContrastiveLossHead`` fed the RAW projections.

Reports the step time of both (CUDA events; forward + backward + AdamW step) and the parity of
loss, encoder-gradient norms and d(log_logit_scale) on identical weights and data.
usage: python tools/config5_step.py [batch] [steps]
"""
import copy
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from transformers import (CLIPTextConfig, CLIPTextModelWithProjection, CLIPVisionConfig,
                          CLIPVisionModelWithProjection)

from mmlearn_b200.task import ContrastiveLossHead

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
torch.manual_seed(0)
vision = CLIPVisionModelWithProjection(CLIPVisionConfig(patch_size=16)).to(dev)
text = CLIPTextModelWithProjection(CLIPTextConfig()).to(dev)
images = torch.randn(batch, 3, 224, 224, device=dev)
tokens = torch.randint(0, 49408, (batch, 77), device=dev)
ids = torch.stack([torch.zeros(batch, dtype=torch.long), torch.arange(batch)], 1).to(dev)


class ReferenceHead(torch.nn.Module):
    """What the reference task + loss do around the towers (contrastive_pretraining.py:447,
    488-489, 607-612; contrastive.py:89-90, 327-340, 135-144), single process."""

    def __init__(self):
        super().__init__()
        self.log_logit_scale = torch.nn.Parameter(torch.ones([]) * math.log(1 / 0.07))

    def forward(self, outputs, example_ids):
        a = F.normalize(F.normalize(outputs["rgb_embedding"], dim=-1), dim=-1)
        b = F.normalize(F.normalize(outputs["text_embedding"], dim=-1), dim=-1)
        with torch.no_grad():
            self.log_logit_scale.clamp_(0, math.log(100.0))
        s = self.log_logit_scale.exp()
        labels = torch.arange(a.shape[0], device=a.device)
        return (F.cross_entropy(s * a @ b.T, labels) + F.cross_entropy(s * b @ a.T, labels)) / 2


def run(head_cls):
    v, t = copy.deepcopy(vision), copy.deepcopy(text)
    head = head_cls().to(dev)
    params = [*v.parameters(), *t.parameters(), *head.parameters()]
    opt = torch.optim.AdamW(params, lr=1e-5)
    ev = []
    first = None
    for i in range(steps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):  # Lightning "bf16-mixed" (cli/run.py:46-49)
            out = {"rgb_embedding": v(pixel_values=images).image_embeds,
                   "text_embedding": t(input_ids=tokens).text_embeds}
            loss = head(out, {"rgb": ids, "text": ids})
        loss.backward()
        if first is None:
            gn = lambda m: float(torch.sqrt(sum((p.grad.float() ** 2).sum() for p in m.parameters()
                                                if p.grad is not None)))
            first = dict(loss=float(loss.detach()), g_vision=gn(v), g_text=gn(t),
                         g_log_scale=float(head.log_logit_scale.grad))
        opt.step()
        e1.record()
        if i >= 2:
            ev.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return first, ms[len(ms) // 2]


ref, ms_ref = run(ReferenceHead)
fused, ms_fused = run(ContrastiveLossHead)
print(f"config 5 (1 GPU, batch {batch}, ViT-B/16 + CLIP text, bf16 autocast, AdamW)")
print(f"  reference loss arithmetic (torch): {ms_ref:8.2f} ms/step   first-step {ref}")
print(f"  fused loss (ContrastiveLossHead):  {ms_fused:8.2f} ms/step   first-step {fused}")
for k in ref:
    rel = abs(fused[k] - ref[k]) / max(abs(ref[k]), 1e-12)
    print(f"  rel diff {k:12s} {rel:.2e}")
