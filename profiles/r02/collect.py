"""Turn the scratch results of profiles/r02/measure_final.sh (gpurun_out/r02/) into the tracked summaries of profiles/r02/.
usage: python profiles/r02/collect.py [gpurun_out/r02]"""
import csv
import gzip
import io
import json
import os
import shutil
import subprocess
import sys
from collections import OrderedDict

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE, "..", "..", "gpurun_out", "r02b")


def bench_lines():
    out = []
    for f in sorted(os.listdir(SRC)):
        if f.startswith("bench_") and f.endswith(".json"):
            for ln in open(os.path.join(SRC, f)):
                ln = ln.strip()
                if ln.startswith("{"):
                    d = json.loads(ln)
                    d["_file"] = f
                    out.append(d)
    with open(os.path.join(HERE, "bench_lines_r02.jsonl"), "w") as fh:
        for d in out:
            fh.write(json.dumps(d) + "\n")
    return out


def short_name(n):
    n = n.split("(")[0]
    return n.replace("void ", "")


def launch_list():
    p = os.path.join(SRC, "launches_cfg2.csv.gz")
    if not os.path.isfile(p):
        return None
    txt = gzip.open(p, "rt").read()
    start = txt.find('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    launches = OrderedDict()
    for r in rows:
        L = launches.setdefault(int(r["ID"]), {"name": short_name(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
        v = float(r["Metric Value"].replace(",", ""))
        u = r["Metric Unit"]
        if r["Metric Name"] == "gpu__time_duration.sum":
            L["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
        else:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
            L[r["Metric Name"]] = v * mult
    ls = list(launches.values())
    shutil.copy(p, os.path.join(HERE, "launches_bench_cfg2_r02.csv.gz"))
    return ls


def summarise_launches(ls, per_step):
    """The bench command runs warm-up steps, the timed step(s), one profiling step and a one-signal parity call; every full
    step is `per_step` launches, so the 4th block of per_step launches is the timed step."""
    step = ls[3 * per_step: 4 * per_step]
    agg = OrderedDict()
    for L in step:
        a = agg.setdefault(L["name"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"] += L.get("us", 0.0)
        a["rd"] += L.get("dram__bytes_read.sum", 0.0)
        a["wr"] += L.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    rd = sum(a["rd"] for a in agg.values())
    wr = sum(a["wr"] for a in agg.values())
    md = ["# r02 - ncu launch list of `python bench.py --steps 1 --warmup 3 --tuning` (cfg2, fp32), one step = %d launches\n" % per_step,
          "`NWCWT_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none`\n(graph replay off under ncu: the same kernels, launched from the host);",
          "per-launch times under ncu are serialised (no overlap between the plan's streams), so the SHARES are what compares with the",
          "bench line's `roofline.single_stream.classes_ms`; the full list is `launches_bench_cfg2_r02.csv.gz`.\n",
          "| kernel | launches | sum us | share | avg us | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        md.append("| `%s` | %d | %.1f | %.1f %% | %.2f | %.1f | %.1f |" % (k[:110], a["n"], a["us"], 100 * a["us"] / tot, a["us"] / a["n"],
                                                                    a["rd"] / 1e6, a["wr"] / 1e6))
    md.append("| **step** | %d | %.1f | | | %.1f | %.1f |\n" % (len(step), tot, rd / 1e6, wr / 1e6))
    open(os.path.join(HERE, "launches_bench_cfg2_summary.md"), "w").write("\n".join(md))
    return tot, rd, wr, len(step)


def ncu_tables():
    for rep, dst in (("ncu_cfg2_rs", "ncu_cfg2_summary.md"), ("ncu_cfg3", "ncu_cfg3_summary.md")):
        raws = [os.path.join(SRC, r + ".raw.csv") for r in ((rep, "ncu_cfg2_ab") if rep == "ncu_cfg2_rs" else (rep,))]
        # (the cfg3 kernel did not change after the first pass of the round: its capture is not repeated)
        raws = [r for r in raws if os.path.isfile(r) and os.path.getsize(r) > 0]
        if not raws:
            continue
        md = subprocess.run([sys.executable, os.path.join(HERE, "summarize_ncu.py")] + raws, stdout=subprocess.PIPE, text=True).stdout
        head = "# r02 - `ncu --set full --clock-control none --import-source on` of the shipped kernels (%s)\n\n" \
               "Commands: `profiles/r02/measure_final.sh`; each capture after the same program had exited 0 without ncu.  " \
               "Replays are cold-cache and serialised.\n\n" % ("cfg2, fp32" if "cfg2" in dst else "cfg3, fp32")
        open(os.path.join(HERE, dst), "w").write(head + md)


if __name__ == "__main__":
    lines = bench_lines()
    print("bench lines:", [(d["_file"], round(d.get("ms_per_step", 0), 3)) for d in lines])
    ls = launch_list()
    if ls:
        main = [d for d in lines if d["_file"] == "bench_cfg2.json"]
        per_step = int(main[0]["gpu_launches"] // main[0]["steps"]) if main else 311
        tot, rd, wr, n = summarise_launches(ls, per_step)
        json.dump({"workload": "cfg2, fp32: one full step (64 signals, %d launches)" % n, "signals": 64,
                   "dram_bytes_per_step": rd + wr, "dram_read": rd, "dram_write": wr,
                   "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --cache-control none --clock-control none over the "
                             "launches of the timed step of `bench.py --steps 1 --warmup 3 --tuning` (profiles/r02/measure_final.sh)",
                   "note": "ncu serialises the launches, so the intermediate (Tm) and the decimated rows (Y) are read back from "
                           "wherever they are when the next kernel starts: nothing is evicted by a concurrent launch group, but "
                           "also nothing overlaps"},
                  open(os.path.join(HERE, "traffic_cfg2.json"), "w"), indent=1)
        print("step under ncu: %.1f us, DRAM %.2f GB read + %.2f GB written" % (tot, rd / 1e9, wr / 1e9))
    ncu_tables()
    for f in ("gputest.log", "smoke.log", "gpu.txt"):
        if os.path.isfile(os.path.join(SRC, f)):
            shutil.copy(os.path.join(SRC, f), os.path.join(HERE, f))
