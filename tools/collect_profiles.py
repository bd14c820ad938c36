"""Turn the raw files of tools/gpu_r2_evidence.sh (gpurun_out/) into the tracked summaries under profiles/ (run here, no GPU):
   python tools/collect_profiles.py [tag]      tag defaults to r02
 * <tag>_launches_step.csv / <tag>_launches_summary.txt : the ncu launch list of three eager steps and its per-kernel totals
 * <tag>_<capture>_ncu.txt                             : key metrics of every .ncu-rep (tools/ncu_key.py)
 * <tag>_step_timeline.csv + <tag>_step_timeline_summary.txt : CUPTI timeline of the captured step, busy time per stream / kernel
 * bench JSON lines, the GPU test log."""
import collections
import csv
import glob
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
src = os.path.join(ROOT, "gpurun_out")
dst = os.path.join(ROOT, "profiles")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("cilrs::", "")
    return re.sub(r"\((?:const |cilrs|__nv|float|long|int|unsigned|void|double).*", "", name)[:70]


def launches():
    path = os.path.join(src, "%s_launches_step.csv" % tag)
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    n = len(rows)
    per = n // 3
    step = rows[2 * per:]          # the third of the three eager steps
    tot = collections.OrderedDict()
    for r in step:
        k = short(r[ki])
        a = tot.setdefault(k, [0, 0.0, 0.0])
        us = float(r[vi].replace(",", "")) / 1e3
        a[0] += 1; a[1] += us; a[2] = max(a[2], us)
    total = sum(a[1] for a in tot.values())
    with open(os.path.join(dst, "%s_launches_summary.txt" % tag), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/prof_step.py 3   (B200, gpurun)\n")
        f.write("# %d launches in three eager training steps at B = 128; this table: the last %d (one step). Per-launch times under ncu are\n" % (n, len(step)))
        f.write("# cold-cache and serialised: compare SHARES with bench.py's roofline.breakdown_ms, not absolutes. Sum = %.1f us.\n" % total)
        for k, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write("%-72s n=%3d total=%8.1f us avg=%6.1f max=%6.1f share=%.3f\n" % (k, a[0], a[1], a[1] / a[0], a[2], a[1] / total))
    shutil.copy(path, os.path.join(dst, "%s_launches_step.csv" % tag))
    print("launch list:", n, "launches,", len(step), "in the summarised step, %.1f us" % total)


def ncu_reps():
    for rep in sorted(glob.glob(os.path.join(src, "%s_prof_*.ncu-rep" % tag))):
        name = os.path.basename(rep)[len(tag) + 6:-8]
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_key.py"), rep], capture_output=True, text=True).stdout
        with open(os.path.join(dst, "%s_%s_ncu.txt" % (tag, name)), "w") as f:
            prog = "python tools/k0_bench.py = K0 alone on 1024 frames 600x800x3" if "k0" in name else "python tools/prof_step.py 3 = three eager steps at B = 128, third step captured"
            f.write("# key metrics of gpurun_out/%s (tools/gpu_r2_evidence.sh; %s)\n" % (os.path.basename(rep), prog))
            f.write(out)
        print(name, len(out.splitlines()), "lines")


def timeline():
    path = os.path.join(src, "%s_step_timeline.csv" % tag)
    if not os.path.exists(path):
        return
    shutil.copy(path, os.path.join(dst, "%s_step_timeline.csv" % tag))
    ev = []
    for r in csv.DictReader(open(path)):
        ev.append((float(r["start_us"]), float(r["dur_us"]), r["stream"], r["name"]))
    # the profile holds three graph replays: take the middle one, delimited by the K0 normalise kernel
    starts = [e[0] for e in ev if "normalize" in e[3] or "preprocess" in e[3]]
    if len(starts) >= 3:
        lo, hi = starts[1], starts[2]
        ev = [e for e in ev if lo <= e[0] < hi]
    span = max(e[0] + e[1] for e in ev) - min(e[0] for e in ev)
    # union of busy intervals (any stream)
    iv = sorted((e[0], e[0] + e[1]) for e in ev)
    busy, cur_s, cur_e = 0.0, None, None
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    per = collections.OrderedDict()
    for s, d, st, n in ev:
        a = per.setdefault(n, [0, 0.0])
        a[0] += 1; a[1] += d
    streams = collections.Counter()
    for s, d, st, n in ev:
        streams[st] += d
    with open(os.path.join(dst, "%s_step_timeline_summary.txt" % tag), "w") as f:
        f.write("# tools/step_timeline.py: CUPTI kernel records of ONE replay of the captured training step (B = 128, 1 GPU)\n")
        f.write("# span %.1f us, some kernel running %.1f us (idle %.1f us), %d kernels; kernel time per stream: %s\n"
                % (span, busy, span - busy, len(ev), ", ".join("%s: %.0f us" % kv for kv in streams.items())))
        for n, a in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write("%-60s n=%3d total=%8.1f us avg=%6.1f\n" % (n[:60], a[0], a[1], a[1] / a[0]))
    print("timeline: span %.1f us busy %.1f us" % (span, busy))


def copies():
    for name in ("%s_bench_1gpu.json" % tag, "%s_bench_reference.json" % tag, "%s_gpu_tests.log" % tag, "%s_k0_bench.json" % tag, "%s_smoke.log" % tag):
        p = os.path.join(src, name)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copy(p, os.path.join(dst, name))
            print("copied", name)


if __name__ == "__main__":
    os.makedirs(dst, exist_ok=True)
    copies()
    launches()
    timeline()
    ncu_reps()
