"""Builds profiles/ncu_r02_summary.md from the round-2 captures brought back in gpurun_out/:
  k2_launches.csv    ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none  python tools/profile_k2.py
  k2prof_r02.ncu-rep ncu --set full --clock-control none --import-source on -k regex:'knn_kernel2|leftover_kernel2|count_psi|
                     layout_kernel|bucket_scatter' --launch-skip 12 --launch-count 6  python tools/profile_k2.py
  g3_launches.csv, g3prof_r02.ncu-rep   the same two passes over tools/profile_g3.py (4-D k-NN entropy, three-level grid)
(read here with `ncu -i ... --page raw/source --csv`; no GPU needed)."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "ncu_r02_summary.md")
LAUNCHES = os.path.join(ROOT, "gpurun_out", "k2_launches.csv")
REP = os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "k2prof_r02.ncu-rep")
G3_LAUNCHES = os.path.join(ROOT, "gpurun_out", "g3_launches.csv")
G3_REP = os.path.join(ROOT, "gpurun_out", "g3prof_r02.ncu-rep")

RAW = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed_pipe_fp64.sum",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def ncu(*args, rep=None):
    return subprocess.run(["ncu", "-i", rep or REP, *args], capture_output=True, text=True).stdout


def g3_launch_list(out):
    if not os.path.exists(G3_LAUNCHES):
        return
    with open(G3_LAUNCHES) as f:
        lines = [l for l in f if not l.startswith("==")]
    d, order = {}, []
    for x in csv.DictReader(lines):
        k = x["ID"]
        if k not in d:
            d[k] = {"name": x["Kernel Name"].replace("k2::<unnamed>::", "k2::")[:70]}
            order.append(k)
        d[k][x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    ids = [k for k in order if "sample_gather" in d[k]["name"]]
    step = order[order.index(ids[-1]) - 1:]
    total = sum(d[k].get("gpu__time_duration.sum", 0) for k in step)
    out.write("## One 4-D k-NN entropy estimate (N = 5 x 10^5, k = 5) through the three-level grid, launch by launch\n\n"
              "`ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none python tools/profile_g3.py` "
              "(second call; the upload and de-interleave precede the first kernel listed).\n\n"
              "| kernel | µs | share | warp instructions (10^6) |\n|---|---|---|---|\n")
    for k in step:
        t = d[k].get("gpu__time_duration.sum", 0)
        out.write(f"| `{d[k]['name']}` | {t / 1000:.1f} | {100 * t / total:.1f} % | {d[k].get('smsp__inst_executed.sum', 0) / 1e6:.2f} |\n")
    out.write(f"| total | {total / 1000:.1f} | | |\n\n")


def launch_list(out):
    with open(LAUNCHES) as f:
        lines = [l for l in f if not l.startswith("==")]
    d, order = {}, []
    for x in csv.DictReader(lines):
        k = x["ID"]
        if k not in d:
            d[k] = {"name": x["Kernel Name"].replace("k2::<unnamed>::", "k2::")[:70]}
            order.append(k)
        d[k][x["Metric Name"]] = float(x["Metric Value"].replace(",", ""))
    ids = [k for k in order if "sample_gather" in d[k]["name"]]
    start = order.index(ids[-2]) - 1          # (two per step: the x column's, then the y column's on the second stream)
    step = order[start:]
    total = sum(d[k].get("gpu__time_duration.sum", 0) for k in step)
    out.write("## One resident step of BASELINE.json configs[1] (N = 10^6, k = 3), launch by launch\n\n"
              "`ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none python tools/profile_k2.py` "
              "(EB2_GRAPH=0; per-launch times are cold-cache and serialised: compare shares).  The y-column grid and the fine "
              "cells overlap the search on the second stream in a normal run.\n\n| kernel | µs | share | warp instructions (10^6) |\n|---|---|---|---|\n")
    for k in step:
        t = d[k].get("gpu__time_duration.sum", 0)
        out.write(f"| `{d[k]['name']}` | {t / 1000:.1f} | {100 * t / total:.1f} % | {d[k].get('smsp__inst_executed.sum', 0) / 1e6:.2f} |\n")
    out.write(f"| total | {total / 1000:.1f} | | |\n\n")


def raw_tables(out, rep=None, title="the pipeline's kernels (N = 10^6)"):
    rows = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv", rep=rep))))
    if len(rows) < 3:
        out.write("(no --set full capture found)\n")
        return
    hdr, units = rows[0], rows[1]
    out.write(f"## `ncu --set full --clock-control none --import-source on` of {title}\n\n")
    names = [r[hdr.index("Kernel Name")].replace("unnamed>::", "")[:48] for r in rows[2:]]
    out.write("| metric | " + " | ".join(f"`{n}`" for n in names) + " |\n|---|" + "---|" * len(names) + "\n")
    for m in RAW:
        if m not in hdr:
            continue
        i = hdr.index(m)
        out.write(f"| {m} [{units[i]}] | " + " | ".join(r[i][:14] for r in rows[2:]) + " |\n")
    out.write("\n")


def source_top(out, kernel, top=16, rep=None):
    rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kernel}", rep=rep))))
    cur, lines = None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) >= 8 and r[0] not in ("", "Line No", "Function Name") and r[2] == "-":
            try:
                lines.append((int(r[7]), int(r[4] or 0), cur, r[0], r[1].strip()[:100]))
            except ValueError:
                pass
    if not lines:
        return
    tot = sum(l[0] for l in lines)
    ts = max(sum(l[1] for l in lines), 1)
    out.write(f"### `{kernel}`: source lines by warp instructions executed (total {tot / 1e6:.1f} x 10^6)\n\n"
              "| instructions | share | stall samples | line |\n|---|---|---|---|\n")
    for l in sorted(lines, reverse=True)[:top]:
        out.write(f"| {l[0]:,} | {100 * l[0] / tot:.1f} % | {100 * l[1] / ts:.1f} % | `{l[2]}:{l[3]}` `{l[4].replace('|', '/')}` |\n")
    out.write("\n")


def main():
    with open(OUT, "w") as out:
        out.write("# ncu summary, round 2 — the bivariate pipeline (`ennemi_b200/csrc/eb2_ksg2.cu`)\n\n")
        launch_list(out)
        raw_tables(out)
        for k in ("knn_kernel2", "count_psi"):
            source_top(out, k)
        if os.path.exists(G3_REP):
            out.write("# The three-level grid (3-D / 4-D k-NN entropy)\n\n")
            g3_launch_list(out)
            raw_tables(out, rep=G3_REP, title="the grid's layout and search kernels (N = 5 x 10^5, 4-D, k = 5)")
            source_top(out, "knn3_kernel", rep=G3_REP)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
