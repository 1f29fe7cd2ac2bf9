"""profiles/dominant_kernel.json from the --set full captures in profiles/r02_final/ (feeds bench.py's roofline.traffic).
The entries carry the content hash of the kernel sources they were captured from; bench.py prints traffic_stale when it differs."""
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
D = os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "r02_final2")
src_hash = importlib.import_module("comfyui-vrgamedevgirl_b200.build")._source_hash()


def summ(tag):
    with open(os.path.join(D, "ncu_%s_summary.json" % tag)) as fh:
        return json.load(fh)


apply_, mom, c1 = summ("apply_f32"), summ("momstore_f32"), summ("configs1_f16")
groups = 16            # bench.py: 128 frames per GPU in groups of 8
per_group = apply_["dram_read_bytes"] + apply_["dram_write_bytes"] + mom["dram_read_bytes"] + mom["dram_write_bytes"]
out = {
    "headline_apply": {
        "what": "bench.py headline step = 16 groups x (k_lab_moments<float,grain,store f-planes> + k_tile<float, colormatch-from-f|lut, unsharp>); "
                "dram bytes of the two kernels of one 8-frame group from their --set full captures, times 16",
        "dram_bytes_per_launch": int(per_group * groups),
        "per_group": {"k_tile<float,20>": {"dram_read": apply_["dram_read_bytes"], "dram_write": apply_["dram_write_bytes"], "duration_us": apply_["duration_us"],
                                           "grid": apply_["grid"], "block": apply_["block"], "dyn_smem_bytes": apply_["dyn_smem_bytes"]},
                      "k_lab_moments<float,1,1>": {"dram_read": mom["dram_read_bytes"], "dram_write": mom["dram_write_bytes"], "duration_us": mom["duration_us"],
                                                  "grid": mom["grid"], "block": mom["block"]}},
        "algorithmic_bytes_per_step": 128 * 2160 * 3840 * 24,
        "capture": "profiles/%s/ncu_apply_f32_* + ncu_momstore_f32_* (tools/r2_run20.sh)" % os.path.basename(D) + "",
        "src_hash": src_hash,
    },
    "configs1_f16": {
        "what": "k_tile<__half, grain|lut, unsharp>, 64 x 1080p fp16 frames, one launch",
        "dram_bytes_per_launch": int(c1["dram_read_bytes"] + c1["dram_write_bytes"]),
        "duration_us": c1["duration_us"], "grid": c1["grid"], "block": c1["block"], "dyn_smem_bytes": c1["dyn_smem_bytes"],
        "algorithmic_bytes_per_launch": 64 * 1080 * 1920 * 12,
        "capture": "profiles/%s/ncu_configs1_f16_* (tools/r2_run20.sh)" % os.path.basename(D) + "",
        "src_hash": src_hash,
    },
}
with open(os.path.join(ROOT, "profiles", "dominant_kernel.json"), "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out, indent=1)[:1200])
