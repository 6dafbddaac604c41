"""Print the headline fields of a bench.py JSON line.  usage: python tools/show_bench.py <log>"""
import json
import sys

for line in open(sys.argv[1]):
    if line.startswith("{"):
        d = json.loads(line)
        e = d.get("e2e", {})
        print(f"{sys.argv[1]}: n_gpus {d.get('n_gpus')} value {d.get('value'):.4g} {d.get('unit')} ms/step {d.get('ms_per_step'):.3f} "
              f"e2e {e.get('value', 0):.4g} graph {d.get('config', {}).get('cuda_graph')} exch {d.get('config', {}).get('grad_exchange')}")
        k = d.get("kernels_ms_per_step")
        if k:
            print("  ", {a: round(b, 3) for a, b in list(k.items())[:12]})
