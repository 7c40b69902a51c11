"""Turn the raw outputs of tools/profile_round.sh (gpurun_out/<round>_*.csv / .ncu-rep) into the tracked summaries under
profiles/ and refresh profiles/igemm_traffic.json (the `traffic` figure bench.py reports).

    python tools/summarize_profile.py r02
"""
import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def rows(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    return list(csv.DictReader(lines))


def val(row):
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v * scale


def per_launch(round_):
    src = ROOT / "gpurun_out" / f"{round_}_igemm_per_launch.csv"
    by_id = collections.OrderedDict()
    for r in rows(src):
        by_id.setdefault(int(r["ID"]), {"grid": r["Grid Size"]})[r["Metric Name"]] = val(r)
    out = [f"# ncu per-launch metrics of the {len(by_id)} igemm launches of ONE forward (1 x 256^2 tile) inside "
           f"`bench.py --steps 20`; serialised, caches flushed by ncu between kernels",
           "# id grid  dram_read_MB dram_write_MB  l2_MB  duration_us  tensor_inst  tensor_pipe_active_%"]
    tot = collections.Counter()
    for i, m in by_id.items():
        rd, wr = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0)
        l2 = m.get("lts__t_bytes.sum", 0)
        du = m.get("gpu__time_duration.sum", 0)
        ti = m.get("sm__inst_executed_pipe_tensor.sum", float("nan"))
        tp = m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", float("nan"))
        tot.update(rd=rd, wr=wr, l2=l2, du=du)
        out.append(f"{i:5d} {m['grid']:>12s} {rd / 1e6:9.3f} {wr / 1e6:9.3f} {l2 / 1e6:9.3f} {du:9.2f} {ti:12.0f} {tp:8.2f}")
    n = len(by_id)
    out.append(f"# totals: dram read {tot['rd'] / 1e6:.1f} MB, dram write {tot['wr'] / 1e6:.1f} MB, L2 traffic "
               f"{tot['l2'] / 1e6:.1f} MB, {tot['du']:.1f} us over {n} launches")
    (ROOT / "profiles" / f"{round_}_igemm_per_launch.txt").write_text("\n".join(out) + "\n")
    traffic = {"dram_bytes_per_launch": (tot["rd"] + tot["wr"]) / n, "launches": n, "workload_tiles": 1,
               "dram_read_bytes_total": tot["rd"], "dram_write_bytes_total": tot["wr"], "l2_bytes_total": tot["l2"],
               "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum (+ lts__t_bytes.sum) over the {n} igemm launches "
                         f"of one 1x256^2 forward (tools/profile_round.sh -> profiles/{round_}_igemm_per_launch.txt). ncu "
                         "flushes caches between replayed kernels: every activation is read from DRAM once; the bf16 "
                         "outputs mostly stay dirty in the 126 MB L2 past the end of the counted window, which is why "
                         "write bytes are small -- lts__t_bytes is the flush-independent figure."}
    (ROOT / "profiles" / "igemm_traffic.json").write_text(json.dumps(traffic, indent=1))
    return traffic


def launches(round_):
    src = ROOT / "gpurun_out" / f"{round_}_launches.csv"
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "summarize_launches.py"), str(src)], capture_output=True,
                         text=True, check=True).stdout
    (ROOT / "profiles" / f"{round_}_launches_bench_steps20.txt").write_text(out)


def full(round_):
    rep = ROOT / "gpurun_out" / f"{round_}_igemm_full.ncu-rep"
    if not rep.exists():
        return
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    keep = ("gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor",
            "sm__pipe_tensor", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "launch__shared_mem_per_block_dynamic")
    rr = list(csv.reader(txt.splitlines()))
    if len(rr) < 3:
        return
    hdr, units = rr[0], rr[1]
    cols = [i for i, h in enumerate(hdr) if any(h.startswith(k) for k in keep)]
    out = [f"# ncu --set full capture of six consecutive igemm launches of one 1x256^2 forward ({rep.name}); selected raw "
           "metrics (tensor-pipe utilisation included), one column per launch"]
    for i in cols:
        out.append(f"{hdr[i]:72s} [{units[i]:>10s}] " + " ".join(f"{r[i]:>12s}" for r in rr[2:]))
    (ROOT / "profiles" / f"{round_}_igemm_ncu_full_summary.txt").write_text("\n".join(out) + "\n")


if __name__ == "__main__":
    r = sys.argv[1] if len(sys.argv) > 1 else "r02"
    launches(r)
    print(per_launch(r))
    full(r)
