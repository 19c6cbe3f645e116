"""gpurun_out/ ncu outputs -> small tracked summaries under profiles/ (run from the repo root)"""
import collections, csv, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r1b"
out = "profiles"
# 1. launch list of the bench command -> copy (kernels of this library only) + per-kernel summary
src = f"gpurun_out/{tag}_launches_bench.csv"
rows = list(csv.reader(open(src)))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
keep = [hdr] + [r for r in rows[start:] if len(r) > mv and "sn::" in r[kn]]
with open(f"{out}/{tag}_launches_bench.csv", "w", newline="") as f:
    csv.writer(f).writerows(keep)
agg = collections.OrderedDict()
for r in keep[1:]:
    agg.setdefault(r[kn], []).append(float(r[mv].replace(",", "")) / 1000.0)
with open(f"{out}/{tag}_launch_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "avg_us", "min_us", "max_us", "note: ncu = cold cache, serialised; forced-mode and device-gated launches of the same kernel are mixed in bench.py"])
    for k, v in agg.items():
        w.writerow([k[:110], len(v), f"{sum(v)/len(v):.1f}", f"{min(v):.1f}", f"{max(v):.1f}"])
# 2. ncu --set full captures -> selected raw metrics
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active"]
REPS = {"r1b": [("gpurun_out/prof_r1_step.ncu-rep", "r1b_ncu_step_kernels.csv"), ("gpurun_out/prof_fsparse_r1.ncu-rep", "r1b_ncu_fwd_sparse.csv"),
                ("gpurun_out/prof_bin_r1.ncu-rep", "r1b_ncu_vox_bin.csv")],
        # third part of the round: mask-driven occupancy forward (1.6 % and empty grids), scanning kernel before it, prepare with the bit mask
        "r1c": [("gpurun_out/prof_fo_r1h.ncu-rep", "r1c_ncu_fwd_occ.csv"), ("gpurun_out/prof_fo0_r1d.ncu-rep", "r1c_ncu_fwd_occ_empty_grids.csv"),
                ("gpurun_out/prof_fs_r1c.ncu-rep", "r1c_ncu_fwd_scan_before.csv"), ("gpurun_out/prof_prep_r1h.ncu-rep", "r1c_ncu_prepare.csv")]}
REPS["r2"] = [("gpurun_out/prof_fwd_r2j.ncu-rep", "r2_ncu_fwd_before_polish.csv"), ("gpurun_out/prof_fwd_r2n.ncu-rep", "r2_ncu_fwd_occ.csv"),
              ("gpurun_out/prof_tapgrad_r2.ncu-rep", "r2_ncu_tapgrad_bits.csv"), ("gpurun_out/r2s_fused.ncu-rep", "r2_ncu_fused_attempt_v4.csv"),
              ("gpurun_out/r2s_ring.ncu-rep", "r2_ncu_ring_attempt.csv"), ("gpurun_out/r2v_vox.ncu-rep", "r2_ncu_voxelize_10M.csv")]
for rep, name in REPS.get(tag, []):
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    cols = [h.index("Kernel Name")] + [h.index(m) for m in WANT if m in h] + [i for i, c in enumerate(h) if "issue_stalled" in c and c.endswith("_per_issue_active.ratio")]
    with open(f"{out}/{name}", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [r[cols[0]][:60] for r in rr[2:]])
        for c in cols[1:]:
            w.writerow([h[c], units[c]] + [r[c] for r in rr[2:]])
if tag == "r2":
    # DRAM bytes per launch for bench.py's roofline.traffic: the forward from this round's capture, the kernels that did not
    # change from round 1's (profiles/r1b_traffic.json)
    import json
    old = json.load(open(f"{out}/r1b_traffic.json"))
    rep = "gpurun_out/prof_fwd_r2n.ncu-rep"
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    def val(row, m):
        i = h.index(m)
        v = float(row[i].replace(",", ""))
        return int(v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]])
    row = rr[-1]
    new = {"source": "ncu --set full --clock-control none, one launch each, config 2 (B=32, 64^3, (9,5,5), float64 boundary); fwd_occ_kernel: "
                     "gpurun_out/prof_fwd_r2n.ncu-rep (scratch/prof_fwd.py, round 2) -> profiles/r2_ncu_fwd_occ.csv; the other kernels are unchanged "
                     "since round 1 (profiles/r1b_traffic.json); dram__bytes_read.sum + dram__bytes_write.sum",
           "fwd_occ_kernel": {"dram_bytes_read": val(row, "dram__bytes_read.sum"), "dram_bytes_write": val(row, "dram__bytes_write.sum"),
                              "algorithmic_bytes": 100663296,
                              "note": "x arrives as one occupancy bit per voxel from the state buffer (1 MB); pred (67 MB float64) is written "
                                      "through the 126 MB L2 and mostly still dirty there when the kernel ends"}}
    for k in ("stencil_fwd_kernel", "g0_kernel", "prepare_f64_kernel"):
        new[k] = old[k]
    # the tap gradient of binary grids (occupancy bits instead of the x tile): this round's capture
    raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_tapgrad_r2.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, units = rr[0], rr[1]
    row = rr[-1]
    new["tapgrad_sparse_kernel"] = {"dram_bytes_read": val(row, "dram__bytes_read.sum"), "dram_bytes_write": val(row, "dram__bytes_write.sum"),
                                    "algorithmic_bytes": 33554432 + 1048576,
                                    "note": "binary grids: G0 (float32) + one occupancy bit per voxel; scratch/dbg_ring.py feeds a G0 tensor that is not L2-resident"}
    json.dump(new, open(f"{out}/r2_traffic.json", "w"), indent=1)
print("ok", os.listdir(out))
