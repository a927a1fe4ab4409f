"""Turns raw ncu output under gpurun_out/ into the tracked summaries under profiles/ (development helper).

  python tests/dev/make_profiles.py launches gpurun_out/launches.csv r1
  python tests/dev/make_profiles.py logmel   gpurun_out/logmel.ncu-rep r1 "<command>"
  python tests/dev/make_profiles.py conv     gpurun_out/conv.ncu-rep   r1 "<command>"
"""
import csv, io, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PROF = os.path.join(ROOT, "profiles")


def short(name):
    m = re.search(r"sedb::(\w+)", name)
    if m:
        return "sedb::" + m.group(1)
    name = re.sub(r"\s+", " ", name).replace(",", ";")
    return "torch: " + name[-60:]


def launches(path, tag):
    text = open(path).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    out = [("id", "kernel", "block", "grid", "gpu__time_duration_ns")]
    agg = {}
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        k = short(r["Kernel Name"])
        out.append((r["ID"], k, r["Block Size"].replace(",", " "), r["Grid Size"].replace(",", " "), int(ns)))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ns
    with open(os.path.join(PROF, f"{tag}_launches.csv"), "w") as f:
        for o in out:
            f.write(",".join(str(x) for x in o) + "\n")
    tot = sum(a[1] for a in agg.values())
    tot_sedb = sum(a[1] for k, a in agg.items() if k.startswith("sedb::"))
    with open(os.path.join(PROF, f"{tag}_launch_summary.csv"), "w") as f:
        f.write("kernel,launches,total_us,share_of_all_pct,share_of_sedb_pct\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{a[0]},{a[1] / 1e3:.1f},{100 * a[1] / tot:.2f},"
                    f"{100 * a[1] / tot_sedb if k.startswith('sedb::') else 0:.2f}\n")
    print(open(os.path.join(PROF, f"{tag}_launch_summary.csv")).read())


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    return [{h: (v, u) for h, u, v in zip(hdr, units, r)} for r in rows[2:]]


def num(d, k, scale_units=True):
    v, u = d[k]
    x = float(v.replace(",", ""))
    if scale_units:
        u = u.split("/")[0]
        x *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-3, "usecond": 1e-3, "ns": 1e-6, "nsecond": 1e-6,
              "second": 1e3, "s": 1e3}.get(u, 1.0)
    return x


def logmel(rep, tag, command):
    d = raw(rep)[0]
    clips = 256
    rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
    sys.path.insert(0, ROOT)
    import bench
    out = {
        "command": command, "kernel": "sedb::logmel_fused_kernel<0>", "clips": clips,
        "source_sha": bench.kernel_source_sha(),
        "gpu_time_ms": num(d, "gpu__time_duration.sum"),
        "dram_bytes_read": rd, "dram_bytes_write": wr, "traffic_bytes_per_launch": rd + wr,
        "algorithmic_bytes_per_launch": clips * 11566592,
        "tensor_pipe_active_pct": num(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": num(d, "sm__inst_executed.sum.pct_of_peak_sustained_elapsed"),
        "ipc": num(d, "sm__inst_executed.avg.per_cycle_elapsed") if "sm__inst_executed.avg.per_cycle_elapsed" in d else None,
        "registers_per_thread": num(d, "launch__registers_per_thread"),
        "dynamic_smem_kb": num(d, "launch__shared_mem_per_block_dynamic") / 1e3,
        "dram_throughput_pct_of_peak": num(d, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "smem_lsu_wavefronts_pct": num(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "smem_tensor_wavefronts_pct": num(d, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        "smem_bank_conflict_wavefronts": num(d, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "local_memory_requests": num(d, "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum") + num(d, "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum"),
        "warp_cycles_per_issued_instruction": num(d, "smsp__average_warp_latency_per_inst_issued.ratio"),
        "stall_long_scoreboard_per_issue": num(d, "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        "stall_barrier_per_issue": num(d, "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
        "stall_short_scoreboard_per_issue": num(d, "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        "stall_mio_throttle_per_issue": num(d, "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"),
    }
    json.dump(out, open(os.path.join(PROF, f"{tag}_logmel_full.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


LAYERS = ["block0.conv2 32->32 182x64 +pool", "block1.conv1 32->64 91x32", "block1.conv2 64->64 91x32 +pool",
          "block2.conv1 64->128 45x16", "block2.conv2 128->128 45x16 +pool", "block3.conv1 128->128 22x8",
          "block3.conv2 128->128 22x8"]


def conv(rep, tag, command):
    ds = raw(rep)
    out = {"command": command, "clips": 256, "layers": []}
    for name, d in zip(LAYERS, ds):
        out["layers"].append({
            "layer": name, "gpu_time_us": num(d, "gpu__time_duration.sum") * 1e3,
            "tensor_pipe_active_pct": num(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
            "smem_tensor_wavefronts_pct": num(d, "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
            "smem_lsu_wavefronts_pct": num(d, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
            "dram_read_mb": num(d, "dram__bytes_read.sum") / 1e6, "dram_write_mb": num(d, "dram__bytes_write.sum") / 1e6,
            "issue_active_pct": num(d, "sm__inst_executed.sum.pct_of_peak_sustained_elapsed"),
            "dynamic_smem_kb": num(d, "launch__shared_mem_per_block_dynamic") / 1e3,
        })
    json.dump(out, open(os.path.join(PROF, f"{tag}_conv_full.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    what, path, tag = sys.argv[1:4]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ""
    {"launches": lambda: launches(path, tag), "logmel": lambda: logmel(path, tag, cmd), "conv": lambda: conv(path, tag, cmd)}[what]()
