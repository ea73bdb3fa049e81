"""Which ATen ops one eagerly issued joint step (c2) still launches, with shapes -- to find avoidable tiny launches."""
import sys, collections, torch
sys.path.insert(0, '.')
import timegan_b200 as tg
from timegan_b200 import train_timegan as tt
dev = torch.device('cuda:0')
torch.manual_seed(0)
m = tg.TimeGAN(14, 64, 64, 3, 0.0).to(dev)
oD = tg.FusedAdam(m.discriminator.parameters(), lr=2e-4, betas=(0.5, 0.9))
oG = tg.FusedAdam(tt._params(m.generator, m.supervisor, m.embedder, m.recovery), lr=1e-3, betas=(0.5, 0.9))
x = torch.rand(256, 768, 14, device=dev)
def step():
    tt.disc_step(m, x, dev, oD, 0.2, 0.3, 0.5, None, 1.0, target_acc=0.525, band=0.15, sync=False)
    tt.gen_step(m, x, dev, oG, 5.0, 0.2, 0.3, 0.5, None, 0.05, 0.05, 64, sync=False)
for _ in range(2): step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=False) as prof:
    step(); torch.cuda.synchronize()
cnt = collections.Counter()
for e in prof.events():
    if e.name in ("aten::fill_", "aten::copy_", "aten::add", "aten::add_", "aten::mul", "aten::zero_", "aten::zeros", "aten::zeros_like", "aten::constant_pad_nd", "aten::contiguous", "aten::clone"):
        cnt[(e.name, str(e.input_shapes)[:80])] += 1
for (n, s), c in cnt.most_common(45):
    print(f"{c:4d} {n:22s} {s}")
