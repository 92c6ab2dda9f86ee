"""ncu CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch of one eager train step,
profiles/run_step.py) -> per C-ABI entry point: launches, total time, DRAM bytes.  bench.py reads the JSON this
writes for `roofline.traffic` (DRAM bytes per launch of the dominant kernel family).

    python profiles/family_traffic.py gpurun_out/step_traffic.csv profiles/r01_family_traffic.json
"""
import collections
import csv
import json
import re
import sys

# kernels launched by each C-ABI entry point (msmc-tts_b200/csrc)
FAMILY = [
    (r"conv_wgrad_umma_kernel|conv_wgrad_reuse_kernel", "msmc_conv_wgrad_umma"),
    (r"conv_umma_reuse_kernel|conv_reuse_persist_kernel", "msmc_conv_forward_umma_reuse"),
    (r"conv_umma_kernel", "msmc_conv_forward_umma"),
    (r"conv_gemm_kernel|conv_direct_small_kernel|conv_c1_kernel", "msmc_conv_forward"),
    (r"conv_wgrad_kernel|conv_wgrad_small_kernel", "msmc_conv_wgrad"),
    (r"wgrad_reduce_kernel", "wgrad_reduce (second pass of both weight-gradient entry points)"),
    (r"weight_image_kernel|weight_image_multi_kernel", "msmc_weight_image(_multi)"),
    (r"weight_norm_fwd_kernel|weight_norm_fwd_multi_kernel", "msmc_weight_norm_fwd(_multi)"),
    (r"weight_norm_bwd_kernel", "msmc_weight_norm_bwd"),
    (r"reflect_fold_kernel", "msmc_reflect_pad_fold"),
    (r"attention_bwd", "msmc_attention_bwd"),
    (r"attention_fwd", "msmc_attention_fwd"),
    (r"add_layernorm_bwd|colsum_partials", "msmc_add_layernorm_bwd"),
    (r"add_layernorm_fwd", "msmc_add_layernorm_fwd"),
    (r"vq_search", "msmc_vq_search"),
    (r"vq_ema", "msmc_vq_ema_update"),
    (r"vq_backward", "msmc_vq_backward"),
    (r"xform_apply", "msmc_xform_apply"),
    (r"adam_multi|adam_prepare", "msmc_adam_multi"),
    (r"l1_multi", "msmc_l1_multi"),
]


def unit_scale(u):
    u = u.strip().lower()
    return {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "nsecond": 1e-3, "us": 1.0,
            "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)


rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
per = collections.defaultdict(dict)
name = {}
for r in rows[1:]:
    try:
        v = float(r[vi].replace(",", "")) * unit_scale(r[ui])
    except ValueError:
        continue
    per[r[ii]][r[mi]] = v
    name[r[ii]] = r[ki]
fam = collections.defaultdict(lambda: {"launches": 0, "us": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0})
for i, m in per.items():
    f = "other (torch glue)"
    for pat, fn in FAMILY:
        if re.search(pat, name[i]):
            f = fn
            break
    d = fam[f]
    d["launches"] += 1
    d["us"] += m.get("gpu__time_duration.sum", 0.0)
    d["dram_read_bytes"] += m.get("dram__bytes_read.sum", 0.0)
    d["dram_write_bytes"] += m.get("dram__bytes_write.sum", 0.0)
tot = sum(d["us"] for d in fam.values())
out = {"source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                 "python profiles/run_step.py (one eager train step, B=16, T=240, K=256)",
       "total_kernel_us": tot, "families": {}}
for f, d in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
    d["share"] = d["us"] / tot
    d["dram_bytes_per_launch"] = (d["dram_read_bytes"] + d["dram_write_bytes"]) / max(1, d["launches"])
    out["families"][f] = d
    print("%9.3f ms %5.1f%% %5d launches  %8.2f MB/launch DRAM  %s" % (
        d["us"] / 1e3, 100 * d["share"], d["launches"], d["dram_bytes_per_launch"] / 1e6, f))
json.dump(out, open(sys.argv[2], "w"), indent=1)
