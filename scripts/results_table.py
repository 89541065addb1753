#!/usr/bin/env python
"""bench.py JSON lines (profiles/r2z_*.json) -> markdown tables of the round's measured numbers."""
import json
import sys


def load(path):
    with open(path) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def main(n1, ref1=None, others=()):
    l = load(n1)
    print("| Record | Gpoint-updates/s | ms / step | of measured HBM peak | launches | SM MHz / throttle |")
    print("|---|---|---|---|---|---|")
    def row(name, r):
        c = r["clocks"]
        extra = f" = {r['timesteps_per_s']:.1f} timesteps/s" if "timesteps_per_s" in r else ""
        print(f"| {name} | {r['value']:.1f}{extra} | {r['ms_per_step']:.4f} | {r['roofline']['frac']:.3f} | {r['gpu_launches']} / {r['steps']} steps | "
              f"{c['sm_mhz']} {','.join(c['reasons']) or '-'} |")
    row("**main line: " + l["config"]["workload"] + "**", l)
    for k, r in (l.get("extra") or {}).items():
        if "value" in r:
            row(k, r)
    e = l.get("e2e")
    if e:
        print(f"\ne2e {e['value']:.2f} Gpoint-updates/s ({e['ms_total']:.0f} ms per {l['steps']}-step job, "
              f"{e['h2d_bytes_per_step'] * l['steps'] / 1e9:.1f} GB up, {e['d2h_bytes_per_step'] * l['steps'] / 1e9:.1f} GB down); "
              f"offload {l['e2e_offload']['value']:.2f}; probe {l['e2e_probe']['value']:.1f}")
    if "cpu_baseline" in l:
        c = l["cpu_baseline"]
        print(f"cpu_baseline {c['value']:.2f} Gpoint-updates/s on {c['cores']} cores ({c['kind']}): {c['sample'][:90]}")
    if ref1:
        r = load(ref1)
        print(f"reference arm {r['value']:.2f} on {r['cpu_baseline']['cores']} cores -> e2e ratio {e['value'] / r['value']:.1f}, device ratio {l['value'] / r['value']:.0f}")
    if "parity" in l:
        print("parity:", l["parity"]["ok"], [c["case"] for c in l["parity"]["cases"]])
    fp = (l.get("extra") or {}).get("fma_parity")
    if fp:
        print("fma_parity:", fp.get("cases"))
    for path in others:
        o = load(path)
        print(f"N={o['n_gpus']}: {o['value']:.1f} Gpoint-updates/s, {o['ms_per_step']:.3f} ms/step, e2e {o.get('e2e', {}).get('value')}, "
              f"parity {(o.get('parity') or {}).get('ok')}, by rank {o['clocks'].get('ms_per_step_by_rank')}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None, sys.argv[3:])
