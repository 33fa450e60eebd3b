"""profiles/roofline_traffic.json from an ncu launch list of one eager micro-batch (tools/collect_evidence.sh):
DRAM bytes (read + write) and duration per launch, averaged per C-ABI entry point, so that bench.py can put the
dominant kernel's measured traffic next to its algorithmic bytes.
usage: python tools/roofline_traffic.py gpurun_out/r02_launches.csv > profiles/roofline_traffic.json"""
import csv
import json
import sys

# kernel function name fragment -> C-ABI entry point whose launches it serves
FAMILIES = [("gemm_tc_kernel", "pb_pw_gemm_tc"), ("wgrad_tc_kernel", "pb_pw_wgrad_tc"),
            ("dw_fwd_tma_kernel", "pb_dwconv3d_fwd + stride-1 pb_dwconv3d_dgrad"), ("dw_s1_mma_kernel", "pb_dwconv3d_fwd/dgrad (mma)"),
            ("dw_fwd3d_tma_kernel", "pb_dwconv3d_fwd (kT,3,3)"), ("dw_dgrad_s2_tma_kernel", "pb_dwconv3d_dgrad (stride 2)"),
            ("dw_wgrad_tma_kernel", "pb_dwconv3d_wgrad"), ("bn_act_fwd_kernel", "pb_bn_act_fwd"),
            ("bn_bwd_reduce_kernel", "pb_bn_act_bwd_reduce"), ("bn_bwd_apply_kernel", "pb_bn_act_bwd_apply"),
            ("bn_bwd_bulk_kernel", "pb_bn_act_bwd_reduce + pb_bn_act_bwd_apply (bulk ring)"),
            ("bn_act_fwd_bulk_kernel", "pb_bn_act_fwd (bulk ring)"),
            ("stem_tc_fwd_kernel", "pb_stem_conv_fwd"), ("stem_tc_wgrad_kernel", "pb_stem_conv_wgrad"),
            ("stem_tma_fwd_kernel", "pb_stem_conv_fwd (TMA)"), ("stem_tma_wgrad_kernel", "pb_stem_conv_wgrad (TMA)"),
            ("colreduce_kernel", "pb_pool_fwd / pb_rowdot"), ("colstats_kernel", "pb_colstats")]


def main(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hdr]
    ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    idi = h.index("ID")
    per = {}
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        d = per.setdefault(r[idi], {"name": r[ki]})
        try:
            d[r[mi]] = float(r[vi].replace(",", ""))
        except ValueError:
            pass
    out, total_ns, n_launch = {}, 0.0, 0
    units = {}
    for r in rows[hdr + 1:]:
        if len(r) > vi:
            units[r[mi]] = r[h.index("Metric Unit")]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0,
             "usecond": 1e3, "msecond": 1e6}
    for d in per.values():
        dur = d.get("gpu__time_duration.sum", 0.0) * scale.get(units.get("gpu__time_duration.sum", "ns"), 1.0)
        rd = d.get("dram__bytes_read.sum", 0.0) * scale.get(units.get("dram__bytes_read.sum", "byte"), 1.0)
        wr = d.get("dram__bytes_write.sum", 0.0) * scale.get(units.get("dram__bytes_write.sum", "byte"), 1.0)
        total_ns += dur
        n_launch += 1
        fam = next((abi for frag, abi in FAMILIES if frag in d["name"]), None)
        if fam is None:
            fam = "other: " + d["name"].split("(")[0].split("<")[0].split("::")[-1][:60]
        o = out.setdefault(fam, {"launches": 0, "ns": 0.0, "dram_read": 0.0, "dram_write": 0.0})
        o["launches"] += 1; o["ns"] += dur; o["dram_read"] += rd; o["dram_write"] += wr
    res = {"_source": path, "_launches": n_launch, "_total_us": total_ns / 1e3,
           "_note": "one eager micro-batch (forward + loss + backward) of bench.py under ncu --profile-from-start off; "
                    "per-launch times are cold-cache and serialised: compare SHARES, not absolutes"}
    for fam, o in sorted(out.items(), key=lambda kv: -kv[1]["ns"]):
        res[fam] = {"launches": o["launches"], "share_of_time": o["ns"] / total_ns if total_ns else 0.0,
                    "avg_us": o["ns"] / o["launches"] / 1e3,
                    "dram_bytes_per_launch": (o["dram_read"] + o["dram_write"]) / o["launches"],
                    "dram_read_per_launch": o["dram_read"] / o["launches"],
                    "dram_write_per_launch": o["dram_write"] / o["launches"]}
    json.dump(res, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
