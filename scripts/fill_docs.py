"""Fills the @PLACEHOLDER@ numbers of DESIGN.md / README.md from a bench line (default profiles/r2_bench_ours.json).
Run once after the final measurement; the placeholders are replaced in place (git keeps the template in history)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_bench_ours.json")
d = json.loads(open(src).read().strip().splitlines()[-1])
r = d["roofline"]
peak = r["peak"]
rows = ["| class (bench name) | kernels | launches | alg. GB / proof | ms | TB/s | of peak |", "|---|---|---:|---:|---:|---:|---:|"]
KERNELS = {
    "polynomialize": "`k_polynomialize` (limbs, E gather from 12 MB of L2-resident subtables, S, lookup outputs)",
    "counters": "`k_cnt_digit_hist/_scan/_starts/_scatter` ×2, `k_cnt_heads`, `k_cnt_finish` (side stream)",
    "eq_build": "`k_eq_split`",
    "mle_dot": "`k_dot_eq<u16/u32/u64>` (claim + 37 openings)",
    "hash_build": "`k_hash_t0`, `k_hash_rw_up_r0` (hashes + tree level 1 + round 0 of the bottom layer), `k_hash_if`",
    "product_tree": "`k_tree_up_r0` (two levels + round 0 of both per launch), `k_tree_tail`, `k_tree_top`",
    "sumcheck_collation": "`k_coll_round` ×3, `k_prod_mid_one`, `k_prod_tail_one`",
    "sumcheck_grand_product": "`k_gp_fold_multi<B,scale>`, `k_gp_fold_multi<X>` ×N, `k_gp_tail` (35 layers per launch)",
    "misc": "`k_gp_coeffs_multi`",
    "ntt": "`k_ntt_cols`, `k_ntt_rows` (weights of the FFT nodes)",
    "gkr_layer_weights": "`k_eq_split_multi`, `k_eq_accumulate`, `k_wiring_runs`, `k_wiring_gather`, `k_concat_items`, `k_ext_split/merge`, `k_dot_wconst`",
    "gkr_layer_sumcheck": "`k_prod_round_multi<FUSE0>`, `k_prod_round_multi` ×N, `k_prod_tail`, `k_copy_items`",
}
tot_ms = tot_gb = tot_l = 0
for k, v in r["per_class"].items():
    tb = v["GBps"] / 1e3
    bold = "**" if k == r["kernel"] else ""
    rows.append(f"| {bold}{k}{bold} | {KERNELS.get(k, '')} | {v['launches']:.0f} | {v['alg_GB']:.3f} | {bold}{v['ms']:.3f}{bold} | {tb:.2f} | {bold}{v['GBps'] / peak:.2f}{bold} |")
    tot_ms += v["ms"]; tot_gb += v["alg_GB"]; tot_l += v["launches"]
rows.append(f"| whole proof, single stream | | {tot_l:.0f} | {tot_gb:.2f} | {tot_ms:.3f} | {tot_gb / tot_ms:.2f} | {tot_gb / tot_ms * 1e3 / peak:.2f} |")
rows.append(f"| whole proof, 4 in flight (the timed region) | | | {tot_gb:.2f} | {d['ms_per_proof']:.3f} | {tot_gb / d['ms_per_proof']:.2f} | {tot_gb / d['ms_per_proof'] * 1e3 / peak:.2f} |")
sub = {
    "@VALUE@": f"{d['value']:.1f}", "@MSPP@": f"{d['ms_per_proof']:.2f}", "@LAUNCHES@": str(d["gpu_launches_per_proof"]),
    "@LAT@": f"{d['single_proof_latency_ms']:.2f}", "@E2E@": f"{d['e2e']['value']:.1f}", "@E2ELAT@": f"{d['e2e']['single_proof_latency_ms']:.2f}",
    "@FRAC@": f"{r['frac']:.2f}", "@CLASS_TABLE@": "\n".join(rows),
}
for name in ("DESIGN.md", "README.md"):
    p = os.path.join(ROOT, name)
    s = open(p).read()
    for a, b in sub.items():
        s = s.replace(a, b)
    open(p, "w").write(s)
print({k: v for k, v in sub.items() if k != "@CLASS_TABLE@"})
