"""Summarise an .ncu-rep (one kernel) into markdown: the metrics BASELINE.json's north_star asks
for (L2/HBM traffic, FP32-pipe and issue-slot utilisation, warp execution efficiency) + hot lines.
usage: ncu_summary.py <report.ncu-rep> <lib.so> <kernel substring> <title>"""
import csv, io, subprocess, sys

rep, lib, kern, title = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u, v = rows[0], rows[1], rows[2]
m = {n: (val, unit) for n, unit, val in zip(h, u, v)}
keys = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock during the capture"),
    ("launch__registers_per_thread", "registers / thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of 64 warps/SM)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency: active lanes per instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue-slot utilisation (%)"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA (FP32) pipe utilisation (%)"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe utilisation (%)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (sqrt/rcp/sin) pipe utilisation (%)"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe utilisation (%)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "FFMA thread-inst / cycle (chip)"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts (% of peak)"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", "shared-memory load bank conflicts"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "shared-memory load wavefronts"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate (%)"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput (% of peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate (%)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput (% of peak)"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__bytes.sum.per_second", "DRAM bandwidth"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "local-memory (traversal stack) load sectors"),
    ("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "local-memory (traversal stack) store sectors"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait (fixed latency)"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall: branch resolving"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard (L1/global/local)"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard (shared/XU)"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall: no instruction (I-cache)"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected"),
]
# the figures bench.py quotes, tied to the source they were measured on (bench.py omits them when csrc/ has changed)
try:
    import glob, hashlib, json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import importlib
    code_sha = importlib.import_module("raytracing-practice_b200.csrc_sha").csrc_sha(root)
    def num(k):
        return float(m[k][0].replace(",", "")) if k in m and m[k][0] != "" else None
    def scaled(k):  # ncu prints byte counts with a unit
        v = num(k)
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(m[k][1], 1) if k in m else 1
        return None if v is None else v * mult
    samples = float(os.environ.get("NCU_SAMPLES", "0"))
    if samples > 0 and os.environ.get("NCU_LATEST", "") == "1":
        json.dump({"csrc_sha": code_sha, "source": os.environ.get("NCU_SOURCE", rep), "kernel": kern, "samples_in_capture": samples,
                   "dram_bytes_per_sample": (scaled("dram__bytes_read.sum") + scaled("dram__bytes_write.sum")) / samples,
                   "issue_slot_utilisation": num("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100.0,
                   "active_lanes_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio"),
                   "alu_pipe_utilisation": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active") / 100.0,
                   "fma_pipe_utilisation": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active") / 100.0,
                   "kernel_ms": scaled("gpu__time_duration.sum") if m.get("gpu__time_duration.sum", ("", ""))[1] == "ms" else num("gpu__time_duration.sum"),
                   "warp_instructions": num("smsp__inst_executed.sum"),
                   "l1_hit_rate": num("l1tex__t_sector_hit_rate.pct"), "local_ld_sectors": num("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum"),
                   "local_st_sectors": num("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum")},
                  open(os.path.join(root, "profiles", "latest_ncu.json"), "w"), indent=1)
except Exception as e:  # noqa: BLE001
    sys.stderr.write(f"latest_ncu.json not written: {e}\n")
print(f"# {title}\n")
print(f"Source: `{rep}` (`ncu --set full --clock-control none --import-source on`), kernel `{kern}`.\n")
print("| metric | value |\n|---|---|")
for k, label in keys:
    if k in m and m[k][0] != "":
        print(f"| {label} (`{k}`) | {m[k][0]} {m[k][1]} |")
print("\n## Hot source lines (PC sampling joined with nvdisasm line info)\n\n```")
sys.stdout.flush()
subprocess.run([sys.executable, __file__.replace("ncu_summary.py", "ncu_by_line.py"), rep, lib, kern, "25"])
print("```")
