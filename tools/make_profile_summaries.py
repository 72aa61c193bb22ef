"""Turn the ncu artefacts in gpurun_out/ into the committed summaries under profiles/."""
import collections, csv, io, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles"); os.makedirs(OUT, exist_ok=True)
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"

def launches(csv_path, md_path, note):
    rows = [r for r in csv.reader(open(csv_path)) if len(r) > 5]
    hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict(); tot = 0.0; n = 0
    for r in rows[1:]:
        if r[idx["Metric Name"]] != "gpu__time_duration.sum": continue
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:70]
        v = float(r[idx["Metric Value"]].replace(",", "")); u = r[idx["Metric Unit"]]
        v = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
        agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += v; tot += v; n += 1
    with open(md_path, "w") as f:
        f.write(f"# ncu launch list ({tag}) — `gpu__time_duration.sum`, `--clock-control none`\n\n{note}\n\n")
        f.write(f"{n} launches captured, {tot:.0f} us of kernel time (cold-cache, serialised: compare SHARES).\n\n")
        f.write("| share | total us | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {100 * v / tot:.1f} % | {v:.1f} | {c} | `{k}` |\n")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active"]

def full(rep, md_path, title, extra=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
    with open(md_path, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on` (values per launch).{extra}\n\n")
        for r in rows[2:]:
            f.write(f"## `{r[hdr.index('Kernel Name')][:110]}`\n\n| metric | value |\n|---|---|\n")
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w); f.write(f"| `{w}` | {r[i]} {rows[1][i]} |\n")
            f.write("\n")
        # hottest SASS lines of the last kernel in the report
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(len(rows) - 3), "--launch-count", "1"],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(io.StringIO(src)))
        secs = []; cur = None
        for r in srows:
            if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; secs.append(cur); continue
            if cur is not None: cur["rows"].append(r)
        if secs:
            h = secs[0]["rows"][0]; ix = {x: i for i, x in enumerate(h)}
            data = [r for r in secs[0]["rows"][1:] if len(r) > 10]
            si, ei = ix["# Samples"], ix["Instructions Executed"]
            tot = sum(int(r[si]) for r in data)
            f.write(f"### hottest SASS lines by warp-stall samples ({tot} samples, {secs[0]['name'][:60]})\n\n| samples | executed | SASS |\n|---:|---:|---|\n")
            for i in sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:16]:
                f.write(f"| {data[i][si]} | {data[i][ei]} | `{data[i][ix['Source']].strip()[:90]}` |\n")

g = os.path.join(ROOT, "gpurun_out")
launches(os.path.join(g, f"launches_{tag}b.csv"), os.path.join(OUT, f"{tag}_launches.md"),
         "Command: `ncu --metrics gpu__time_duration.sum --clock-control none -s 170 -c 140 --csv python bench.py --steps 2 --warmup 3 --no-cpu` "
         "(bf16 tier, 10 k molecules; the window covers about two fwd+bwd steps).")
full(os.path.join(g, f"prof_fused_{tag}.ncu-rep"), os.path.join(OUT, f"{tag}_fused_step_ncu.md"), f"Fused depth-step kernel ({tag})",
     " Template arguments <ACT, FIRST, HAS_BIAS, MODE>: launches in order = forward first step (`FIRST`, reads H_0 only, also stores M^1), "
     "forward t>=2 step (the roofline kernel; algorithmic bytes at this size: 910.6 MB), backward mirror step with tau' mask + G output "
     "(MODE 1), last backward mirror step with the tau'(H_0) mask (MODE 3).")
full(os.path.join(g, f"prof_gemm_{tag}.ncu-rep"), os.path.join(OUT, f"{tag}_gemm_ncu.md"), f"tcgen05 linear / weight-gradient kernels ({tag})")

# DRAM traffic of the forward t>=2 launch -> profiles/fused_step_traffic.json (bench.py reports it as roofline.traffic)
import json
raw = subprocess.run(["ncu", "-i", os.path.join(g, f"prof_fused_{tag}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr = rows[0]
def num(r, name):
    i = hdr.index(name); v = float(r[i].replace(",", "")); u = rows[1][i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    m = re.search(r"k_bond_step_fused<([^>]*)>", name)
    targs = [int(re.sub(r"[^0-9]", "", a.replace("true", "1").replace("false", "0")) or 0) for a in m.group(1).split(",")] if m else []
    if len(targs) == 4 and targs[1] == 0 and targs[3] == 0:      # <ACT, FIRST=0, HAS_BIAS, MODE=0>: forward, t >= 2
        rd, wr = num(r, "dram__bytes_read.sum"), num(r, "dram__bytes_write.sum")
        json.dump({"kernel": "k_bond_step_fused<RELU, t>=2, forward>", "directed_edges": 502000, "atoms": 249437, "precision": "bf16",
                   "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
                   "algorithmic_bytes": 910621748, "source": f"profiles/{tag}_fused_step_ncu.md (ncu --set full, one launch)"},
                  open(os.path.join(OUT, "fused_step_traffic.json"), "w"), indent=1)
        print("traffic:", rd + wr)
        break
else:
    print("WARNING: forward t>=2 launch not found in the fused report; kernels:", [r[hdr.index("Kernel Name")][:80] for r in rows[2:]])
print(open(os.path.join(OUT, f"{tag}_launches.md")).read()[:2500])
