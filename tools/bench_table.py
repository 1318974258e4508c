#!/usr/bin/env python
"""Markdown per-kernel table of a bench.py JSON line (for profiles/README.md).

    python tools/bench_table.py profiles/r2_bench.json [profiles/traffic.json]
"""
import json
import sys


def main():
    d = None
    for line in open(sys.argv[1]):
        if line.startswith("{"):
            d = json.loads(line)
    traffic = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else {}
    r = d["roofline"]
    print("| Kernel | ms per step | design GB per step | achieved GB/s | of HBM peak | ncu DRAM GB per launch |")
    print("|---|---|---|---|---|---|")
    for k, v in r["kernels"].items():
        b = v.get("design_bytes_per_step")
        t = traffic.get(k.split("(")[0]) or traffic.get(k)
        print(f"| `{k}` | {v['ms_per_step']:.2f} | {b / 1e9:.2f} | {v['achieved_gbs']:.0f} | {100 * v['frac']:.1f} % | {t / 1e9:.2f} |" if b else
              f"| `{k}` | {v['ms_per_step']:.2f} | – | – | – | {(t or 0) / 1e9:.2f} |")
    print()
    s8 = r.get("survey_8d", {})
    for k, v in s8.items():
        print(f"* SURVEY §8(d) `{k}`: {v['bytes_per_step'] / 1e9:.2f} GB per step in {v['ms_per_step']:.2f} ms = {v['achieved_gbs']:.0f} GB/s = "
              f"{100 * v['frac']:.1f} % of the measured HBM peak ({r['peak']:.0f} GB/s)")
    op = r.get("other_pipes", {})
    if op:
        print(f"* matcher: {op['match_distance_pairs_per_s']:.3g} 256-bit distances/s = {100 * op['match_frac_of_popc_peak']:.0f} % of the measured popc.b64 peak")
    e = d["e2e"]
    print(f"* step: {d['ms_per_step']:.2f} ms device-resident ({d['value']:.0f} {d['unit']}), {e['ms_per_step']:.2f} ms end to end "
          f"({e['value']:.0f} {e['unit']}; {e['h2d_bytes_per_step'] / 1e6:.0f} MB in, {e['d2h_bytes_per_step'] / 1e6:.0f} MB out per step); "
          f"clocks {d['clocks']['sm_mhz']:.0f} MHz, reasons {d['clocks']['reasons']}; {d['gpu_launches']} kernel launches in the timed region")
    cb = d.get("cpu_baseline")
    if cb:
        print(f"* cpu_baseline: {cb['value']:.2f} {cb['unit']} on {cb['cores']} thread(s) ({cb['kind']}; {cb['sample']})")
    for name, x in (d.get("extra_configs") or {}).items():
        if "value" in x:
            print(f"* `{name}`: {x['value']:.4g} {x['unit']} ({x['ms_per_step']:.3f} ms per step), e2e {x['e2e']['value']:.4g}")


if __name__ == "__main__":
    main()
