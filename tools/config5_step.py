"""BASELINE.json config 5: a ContrastivePretraining-style training step with HuggingFace CLIP towers
(ViT-B/16 vision + CLIP text transformer, random init, synthetic image/text batch, bf16 autocast,
AdamW), on 1 GPU or -- under torchrun -- data-parallel on N GPUs (DistributedDataParallel around
both towers, as Lightning's DDP strategy wraps the task).  Two arms on identical weights and data:

* reference: the reference's loss arithmetic in plain torch -- F.normalize (task, then loss:
  contrastive_pretraining.py:447, contrastive.py:89-90), autograd all-gather of both modalities
  (contrastive.py:506), ``logit_scale * local @ all.T`` both ways (contrastive.py:327-332), labels
  offset by the rank's rows (:196-212), two cross-entropies (:135-144).  Its id matching
  (two host syncs per pair and step) is left out, which flatters this arm;
* fused: ``mmlearn_b200.task.ContrastiveLossHead`` fed the RAW projections.

Reports the median step time of both (CUDA events; forward + backward incl. DDP's gradient
all-reduce + AdamW step; max over ranks) and the parity of the loss trajectory, the encoder
gradient norms and d(log_logit_scale) over the steps.
usage: [torchrun --nproc-per-node N] python tools/config5_step.py [batch per GPU] [steps]
"""
import copy
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import torch.nn.functional as F
from transformers import (CLIPTextConfig, CLIPTextModelWithProjection, CLIPVisionConfig,
                          CLIPVisionModelWithProjection)

from mmlearn_b200.task import ContrastiveLossHead

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)  # identical initial weights on every rank
vision = CLIPVisionModelWithProjection(CLIPVisionConfig(patch_size=16)).to(dev)
text = CLIPTextModelWithProjection(CLIPTextConfig()).to(dev)
g = torch.Generator(device=dev).manual_seed(100 + rank)  # every rank its own data shard
images = torch.randn(batch, 3, 224, 224, device=dev, generator=g)
tokens = torch.randint(0, 49408, (batch, 77), device=dev, generator=g)
ids = torch.stack([torch.zeros(batch, dtype=torch.long),
                   torch.arange(rank * batch, (rank + 1) * batch)], 1).to(dev)


class ReferenceHead(torch.nn.Module):
    """What the reference task + loss do around the towers (see the module docstring)."""

    def __init__(self):
        super().__init__()
        self.log_logit_scale = torch.nn.Parameter(torch.ones([]) * math.log(1 / 0.07))

    def forward(self, outputs, example_ids):
        a = F.normalize(F.normalize(outputs["rgb_embedding"], dim=-1), dim=-1)
        b = F.normalize(F.normalize(outputs["text_embedding"], dim=-1), dim=-1)
        with torch.no_grad():
            self.log_logit_scale.clamp_(0, math.log(100.0))
        s = self.log_logit_scale.exp()
        if world > 1:
            import torch.distributed.nn.functional as dfn

            all_a = torch.cat(dfn.all_gather(a), 0)
            all_b = torch.cat(dfn.all_gather(b), 0)
        else:
            all_a, all_b = a, b
        labels = torch.arange(a.shape[0], device=a.device) + rank * a.shape[0]
        return (F.cross_entropy(s * a @ all_b.T, labels) + F.cross_entropy(s * b @ all_a.T, labels)) / 2


def run(head_cls):
    v, t = copy.deepcopy(vision), copy.deepcopy(text)
    head = head_cls().to(dev)
    if world > 1:
        from torch.nn.parallel import DistributedDataParallel as DDP

        v = DDP(v, device_ids=[local_rank])
        t = DDP(t, device_ids=[local_rank])
        if isinstance(head, ReferenceHead):  # noqa: SIM102
            head = DDP(head, device_ids=[local_rank])  # the scale is a task parameter under DDP
    params = [*v.parameters(), *t.parameters(), *head.parameters()]
    opt = torch.optim.AdamW(params, lr=1e-5)
    gn = lambda m: float(torch.sqrt(sum((p.grad.float() ** 2).sum() for p in m.parameters()
                                        if p.grad is not None)))
    ev, hist = [], []
    for i in range(steps + 2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):  # Lightning "bf16-mixed" (cli/run.py:46-49)
            out = {"rgb_embedding": v(pixel_values=images).image_embeds,
                   "text_embedding": t(input_ids=tokens).text_embeds}
            loss = head(out, {"rgb": ids, "text": ids})
        loss.backward()
        hd = head.module if hasattr(head, "module") else head
        gls = hd.log_logit_scale.grad
        if world > 1 and not isinstance(hd, ReferenceHead):
            # the fused head is not DDP-wrapped: average its scale gradient like DDP would
            dist.all_reduce(gls)
            gls /= world
        opt.step()
        e1.record()
        hist.append(dict(loss=loss.detach().float(), g_vision=gn(v), g_text=gn(t),
                         g_log_scale=float(gls)))
        if i >= 2:
            ev.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    med = torch.tensor(ms[len(ms) // 2], device=dev)
    if world > 1:
        dist.all_reduce(med, op=dist.ReduceOp.MAX)
    for h in hist:
        h["loss"] = float(h["loss"])
    return hist, float(med)


def deferred_head():
    head = ContrastiveLossHead()
    head.loss_fn.deferred_verification = True  # no host wait inside the loss forward
    return head


ref, ms_ref = run(ReferenceHead)
fused, ms_fused = run(ContrastiveLossHead)
_, ms_deferred = run(deferred_head)
if rank == 0:
    print(f"config 5 ({world} GPU, batch {batch} per GPU = {batch * world} global, ViT-B/16 + CLIP "
          f"text, bf16 autocast, AdamW, {'DDP' if world > 1 else 'single process'})")
    print(f"  reference loss arithmetic (torch): {ms_ref:8.2f} ms/step   first-step {ref[0]}")
    print(f"  fused loss (ContrastiveLossHead):  {ms_fused:8.2f} ms/step   first-step {fused[0]}")
    print(f"  fused, deferred verification:      {ms_deferred:8.2f} ms/step")
    for k in ref[0]:
        rel = [abs(f[k] - r[k]) / max(abs(r[k]), 1e-12) for f, r in zip(fused, ref)]
        print(f"  rel diff {k:12s} first step {rel[0]:.2e}   max over {len(rel)} steps {max(rel):.2e}")
    print("  rank-0 loss by step, reference:", " ".join(f"{h['loss']:.4f}" for h in ref))
    print("  rank-0 loss by step, fused:    ", " ".join(f"{h['loss']:.4f}" for h in fused))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
