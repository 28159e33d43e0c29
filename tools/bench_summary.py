"""Short summary of a bench.py JSON line: python tools/bench_summary.py gpurun_out/x.json"""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print(f"{f}: value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  N={d['n_gpus']}  clocks {d.get('clocks')}")
    r = d.get("roofline") or {}
    if r:
        print(f"  dominant {r.get('kernel')} {r.get('achieved', 0):.0f} GB/s frac {r.get('frac', 0):.3f}  chain {r.get('chain')}")
        for k, v in (r.get("kernels") or {}).items():
            print(f"    {k:30s} {v['ms_per_launch']:.3f} ms x{v['launches_per_step']:.0f}  {v['gbs']:.0f} GB/s  {v['gbs'] / r['peak']:.3f}")
    e = d.get("e2e") or {}
    if e:
        print(f"  e2e {e['value']:.0f}  frac_of_copy_ceiling {e.get('frac_of_copy_ceiling')}  "
              + "  ".join(f"{k}: {v.get('value', v.get('audio_s_per_s')):.0f}" for k, v in e.items() if isinstance(v, dict) and k != 'copy_ceiling'))
    if "check" in d and "gate" in d["check"]:
        print("  gate", {k: v for k, v in d["check"]["gate"].items() if k not in ("what", "tol")})
    for k, v in (d.get("extra") or {}).items():
        ch = ((v.get("roofline") or {}).get("chain") or {}).get("frac")
        print(f"  extra.{k}: {v.get('value', 0):.0f}  ms/step {v.get('ms_per_step', 0):.2f}  chain frac {ch}  gate {((v.get('check') or {}).get('gate') or {}).get('pass')}")
