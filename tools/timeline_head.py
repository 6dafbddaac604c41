"""Development tool: SM-clock timeline of the gen-2 radiance-head forward kernel (CTA 0 only).
Needs the instrumented build:  bash tools/build_timeline_lib.sh
  SPF_LIBRARY=spurfies_b200/csrc/libspurfies_b200_tl.so python tools/timeline_head.py"""
import ctypes as C, sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from spurfies_b200 import _lib, fields
from spurfies_b200.fields import SlotSet, RadianceHead

dev = torch.device('cuda', 0)
sc, model = bench.build_scene(dev, precision='bf16')
g = torch.Generator().manual_seed(0)
R, S = 4096, 80
n = R * S
q = (sc["pts"][torch.randint(0, 100000, (n,), generator=g)] + 0.03 * torch.randn(n, 3, generator=g)).cuda().contiguous()
slots = SlotSet(model._grid().query_points(q, 8, 2.0))
fields.set_precision("bf16")
fc = [m for m in model.F_color if isinstance(m, torch.nn.Linear)]
rl = [m for m in model.R if isinstance(m, torch.nn.Linear)]
prm = [fc[3].weight, fc[3].bias, rl[0].weight, rl[0].bias, rl[1].weight, rl[1].bias, rl[2].weight, rl[2].bias]
dirs = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).cuda().contiguous()
hbar = (0.5 * torch.randn(n, 256, generator=g)).cuda().requires_grad_()
buf = (C.c_ulonglong * (4 * 8192))()
NAMES = {0: "E.ready(acc+drain)", 1: "E.compute_done", 2: "E.signalled", 3: "E.gather_done", 4: "E.gather_signalled",
         5: "E.iter_end", 6: "E.acc_seen", 10: "M.a_ready_seen", 11: "M.issued"}
for i in range(3):
    torch.cuda.synchronize()
    _lib.lib.spf_debug_timeline(buf, 8192)
    _lib.profile_reset(True)
    rgb = RadianceHead.apply(hbar, *prm, dirs, slots, S)
    torch.cuda.synchronize()
    pr = _lib.profile_collect()
    print("V", slots.V, {k: round(v["ms"], 4) for k, v in pr.items()})
nev = _lib.lib.spf_debug_timeline(buf, 8192)
ev = sorted([(buf[4*i+3], buf[4*i], buf[4*i+1], buf[4*i+2]) for i in range(nev)])
t0 = ev[0][0]
print("events", nev, "span cycles", ev[-1][0] - t0)
last = {}
dur = collections.defaultdict(list)
for c, e, t, l in ev:
    if e >= 10:
        continue
    if t in last:
        pe, pc = last[t]
        dur[(NAMES.get(pe, pe), NAMES.get(e, e))].append(c - pc)
    last[t] = (e, c)
print("phase (from -> to): count, mean cycles, total cycles")
for k, v in sorted(dur.items(), key=lambda kv: -sum(kv[1])):
    print(f"  {k[0]:>22s} -> {k[1]:<22s} {len(v):5d} {sum(v)/len(v):9.0f} {sum(v):10d}")
for c, e, t, l in ev[:160]:
    print(f"{c - t0:9d} cyc  {'  ' if t == 0 else '                          '}tile{t} L{l} {NAMES.get(e, e)}")
