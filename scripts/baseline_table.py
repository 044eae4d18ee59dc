"""Prints BASELINE.md section 2's table rows from the bench lines under profiles/ (one JSON line per file)."""
import json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ROWS = [("C1 64×48 wall 0.20 m", "Simple", "r2_final_bench_C1.json"),
        ("C2 640×480 sphere room 0.10 m, truncation 0.4 m", "Merged", "r2_final_bench_C2.json"),
        ("C2 with the literal \"4 m truncation\"", "Merged", "r2_final_bench_C2_trunc4.json"),
        ("C3 640×480 room sequence 0.05 m", "Fast", "r2_final_bench_C3.json"),
        ("C4 640×480 0.05 m + ESDF update every scan (ROS defaults)", "Merged + ESDF", "r2_final_bench_C4.json"),
        ("bench (metric config): 640×480 room 0.05 m", "Merged", "r2_final_bench.json"),
        ("C5 2048×128 LiDAR 0.05 m (12 M updates/scan)", "Merged", "r2_final_bench_C5.json")]


def fmt(v):
    return f"{v / 1e6:.4g} M"


for name, integ, fn in ROWS:
    path = os.path.join(ROOT, "profiles", fn)
    if not os.path.exists(path):
        print(f"| {name} | {integ} | (missing {fn}) |")
        continue
    d = json.loads(open(path).read().strip().splitlines()[-1])
    cb = d["cpu_baseline"]
    m = re.search(r"fastest of \{([^}]*)\}", cb["sample"])
    times = {}
    if m:
        for part in m.group(1).split(","):
            k, v = part.split(":")
            times[k.strip().strip("'")] = v.strip()
    order = ["1", "4", "8", "16", "32"]
    rest = [k for k in times if k not in order]
    ms = "/".join(f"{float(times[k]):.1f}" for k in order + rest if k in times and re.match(r"^[0-9.]+$", times[k]))
    sync = d["synchronous_call"]
    e2e = d["e2e"]
    print(f"| {name} | {integ} | {cb['cores']} ({ms}) | {fmt(cb['value'])} | {fmt(cb['o3_x86_64_v3']['value'])} | "
          f"{fmt(d['value'])} ({d['ms_per_step']:.3f}) | {fmt(sync['value'])} ({sync['ms_per_step']:.3f} ms) | {fmt(e2e['value'])} | "
          f"{e2e['value'] / cb['value']:.0f}× | {100 * d['roofline']['frac']:.2f} % | `profiles/{fn}` |")
