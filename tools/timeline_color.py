"""Development tool: SM-clock timeline of the gen-2 colour forward / backward kernels (CTA 0 only).
Needs the instrumented build:  bash tools/build_timeline_lib.sh
  SPF_LIBRARY=spurfies_b200/csrc/libspurfies_b200_tl.so python tools/timeline_color.py [fwd|bwd] [debug_mode]"""
import ctypes as C, sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from spurfies_b200 import _lib, fields
from spurfies_b200.fields import SlotSet, ColorField

which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = torch.device('cuda', 0)
sc, model = bench.build_scene(dev, precision='bf16')
g = torch.Generator().manual_seed(0)
n = 160000
q = (sc["pts"][torch.randint(0, 100000, (n,), generator=g)] + 0.01 * torch.randn(n, 3, generator=g)).cuda().contiguous()
slots = SlotSet(model._grid().query_points(q, 8, 2.0))
fields.set_precision("bf16")
fc = [m for m in model.F_color if isinstance(m, torch.nn.Linear)]
buf = (C.c_ulonglong * (4 * 8192))()
_lib.lib.spf_debug_mode(mode)
NAMES = {0: "E.ready(acc+drain)", 1: "E.compute_done", 2: "E.signalled", 3: "E.gather_done", 4: "E.bar_passed", 4: "E.gather_signalled",
         5: "E.iter_end", 6: "E.acc_seen", 10: "M.a_ready_seen", 11: "M.issued"}


def run(backward):
    hbar = ColorField.apply(model.neural_feats_color, fc[0].weight, fc[0].bias, fc[1].weight, fc[1].bias, fc[2].weight,
                            fc[2].bias, q, slots, model.neural_pts, 45.0)
    if backward:
        torch.cuda.synchronize()
        _lib.lib.spf_debug_timeline(buf, 8192)   # drop the forward's events
        d = torch.ones_like(hbar)
        if os.environ.get("SPF_TL_COMPACT", "1") == "1":   # the training path: gradient arrives as compact bf16 tiles
            from spurfies_b200.fields import Arena
            dc = Arena.get("tl.d_hbc", (slots.rows_alloc(1), 256), torch.float16, d.device)
            dc.fill_(1.0)
            slots.d_hb_compact = (dc, d.data_ptr(), fields.grad_scale(d, target=1.0))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        hbar.backward(d)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    return None


for i in range(3):
    torch.cuda.synchronize()
    _lib.lib.spf_debug_timeline(buf, 8192)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ms_b = run(which == "bwd")
    e1.record()
    torch.cuda.synchronize()
    print("V", slots.V, "pairs", slots.V * 8, "ms", ms_b if ms_b is not None else e0.elapsed_time(e1), "(incl. packing / wgrad launches)")
nev = _lib.lib.spf_debug_timeline(buf, 8192)
ev = sorted([(buf[4*i+3], buf[4*i], buf[4*i+1], buf[4*i+2]) for i in range(nev)])
# the backward pass also launches wgrad kernels etc: keep only CTA-0 events of the colour kernel (all we record)
t0 = ev[0][0]
print("events", nev, "span cycles", ev[-1][0] - t0)
# phase statistics per tile group: time between consecutive events of the same epilogue group
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
mm = [(c, e, t, l) for c, e, t, l in ev if e >= 10]
gaps = collections.defaultdict(list)
prev = None
for c, e, t, l in mm:
    if prev is not None:
        gaps[(NAMES[prev[1]], NAMES[e])].append(c - prev[0])
    prev = (c, e, t, l)
print("MMA warp (from -> to): count, mean cycles")
for k, v in gaps.items():
    print(f"  {k[0]:>16s} -> {k[1]:<16s} {len(v):5d} {sum(v)/len(v):9.0f}")
for c, e, t, l in ev[200:330]:
    print(f"{c - t0:9d} cyc  {'  ' if t == 0 else '                          '}tile{t} L{l} {NAMES.get(e, e)}")
