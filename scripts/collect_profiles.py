"""Turn the outputs of scripts/gpu_final.sh (gpurun_out/fin_*) and scripts/gpu_scale.sh (gpurun_out/scale_*.json) into the
tracked files under profiles/ (round-2 names) and print the key numbers for profiles/README.md."""
import json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
O, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def last_json(path):
    try:
        lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
        return json.loads(lines[-1])
    except Exception as e:
        print("!!", path, e)
        return None


def sh(*cmd):
    return subprocess.run(list(cmd), capture_output=True, text=True, cwd=ROOT).stdout


benches = {"bench_cfg3": "r02_bench_cfg3_1gpu.json", "reference": "r02_bench_reference_cpu.json",
           "bench_cfg3_3xtf32": "r02_bench_cfg3_1gpu_3xtf32.json", "bench_cfg3_b1": "r02_bench_cfg3_b1_per_gpu_load.json",
           "bench_cfg2": "r02_bench_cfg2_1gpu.json", "bench_cfg4": "r02_bench_cfg4_1gpu.json", "bench_cfg5": "r02_bench_cfg5_1gpu.json"}
for src, dst in benches.items():
    d = last_json(os.path.join(O, f"fin_{src}.log"))
    if d is None:
        continue
    json.dump(d, open(os.path.join(P, dst), "w"), indent=1)
    e2e = d.get("e2e", {})
    print(f"{dst}: {d['value']} {d['unit']}, {d['ms_per_step']} ms/step, e2e {e2e.get('value')} (u8 {e2e.get('u8', {}).get('value')}), "
          f"roofline {d.get('roofline', {}).get('achieved')} {d.get('roofline', {}).get('unit')} frac {d.get('roofline', {}).get('frac')}, "
          f"vq {d.get('vq_lookup')}, cpu {d.get('cpu_baseline', {}).get('value')}, parity {d.get('parity')}, clocks {d.get('clocks')}")
for n in (1, 2, 4, 8):
    d = last_json(os.path.join(O, f"scale_{n}.json"))
    if d:
        json.dump(d, open(os.path.join(P, f"r02_scale_n{n}.json"), "w"), indent=1)
        print(f"scale n={n}: {d['value']} frames/s {d['ms_per_step']} ms e2e {d['e2e']['value']} gathered==single {d.get('gathered_codes_equal_single_gpu')}")
for tag, name in (("launches", "r02_launches_cfg3"), ("launches_b1", "r02_launches_cfg3_b1")):
    src = os.path.join(O, f"fin_{tag}.csv")
    if os.path.exists(src):
        shutil.copy(src, os.path.join(P, name + ".csv"))
        open(os.path.join(P, name + ".md"), "w").write(sh(sys.executable, "scripts/launch_summary.py", src).replace(src, f"profiles/{name}.csv"))
for tag, name, top in (("attn_f16", "r02_ncu_attn_f16", 20), ("gemm_ff1", "r02_ncu_gemm_ff1", 12), ("peg4", "r02_ncu_peg", 12),
                       ("vq", "r02_ncu_vq", 12), ("ln", "r02_ncu_layernorm", 8)):
    rep = os.path.join(O, f"fin_full_{tag}.ncu-rep")
    if os.path.exists(rep):
        body = sh(sys.executable, "scripts/ncu_summary.py", rep) + "\n```\n" + sh(sys.executable, "scripts/ncu_hot.py", rep, str(top)) + "```\n"
        open(os.path.join(P, name + ".md"), "w").write(body.replace(rep, f"gpurun_out/fin_full_{tag}.ncu-rep (scripts/gpu_final.sh)"))
        print("wrote", name)
if os.path.exists(os.path.join(O, "fin_gemm_shapes.log")):
    shutil.copy(os.path.join(O, "fin_gemm_shapes.log"), os.path.join(P, "r02_gemm_shapes_final.txt"))
