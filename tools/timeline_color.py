"""Development tool: SM-clock timeline of the gen-2 colour forward kernel (library built with SPF_TIMELINE=1)."""
import ctypes as C, sys
sys.path.insert(0, '/root/repo')
import torch
import bench
from spurfies_b200 import _lib, fields
from spurfies_b200.fields import SlotSet, ColorField
dev = torch.device('cuda', 0)
sc, model = bench.build_scene(dev, precision='bf16')
g = torch.Generator().manual_seed(0)
n = 160000
q = (sc["pts"][torch.randint(0, 100000, (n,), generator=g)] + 0.01 * torch.randn(n, 3, generator=g)).cuda().contiguous()
slots = SlotSet(model._grid().query_points(q, 8, 2.0))
fields.set_precision("bf16")
fc = [m for m in model.F_color if isinstance(m, torch.nn.Linear)]
buf = (C.c_ulonglong * (4 * 8192))()
_lib.lib.spf_debug_mode(0)
for i in range(3):
    torch.cuda.synchronize()
    _lib.lib.spf_debug_timeline(buf, 8192)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hbar = ColorField.apply(model.neural_feats_color, fc[0].weight, fc[0].bias, fc[1].weight, fc[1].bias, fc[2].weight,
                            fc[2].bias, q, slots, model.neural_pts, 45.0)
    e1.record()
    torch.cuda.synchronize()
    print("V", slots.V, "pairs", slots.V * 8, "ms (incl. weight packing launches)", e0.elapsed_time(e1))
nev = _lib.lib.spf_debug_timeline(buf, 8192)
ev = sorted([(buf[4*i+3], buf[4*i], buf[4*i+1], buf[4*i+2]) for i in range(nev)])
t0 = ev[0][0]
names = {0: "E.acc_seen", 1: "E.compute_done", 2: "E.signalled", 3: "E.gather_done", 4: "E.gather_signalled", 5: "E.iter_end", 10: "M.a_ready_seen", 11: "M.issued"}
print("events", nev)
for c, e, t, l in ev[:150]:
    print(f"{c - t0:9d} cyc  {'  ' if t == 0 else '                          '}tile{t} L{l} {names.get(e, e)}")
