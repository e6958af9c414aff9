#!/usr/bin/env python3
"""Builds differently tuned variants of the library for A/B runs of K1 on the GPU box (ring depth, tile size, CTAs per SM).
    python tools/k1_variants.py            # build/variants/librv_<name>.so
The GPU side: RV_LIBRARY_PATH=build/variants/librv_<name>.so python bench.py --no-e2e --no-cpu-baseline --no-rows ..."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repas_vision_b200 import _build  # noqa: E402

VARIANTS = {
    "s2_i8_c4": "-DRV_K1_STAGES=2 -DRV_K1_ITERS=8 -DRV_K1_WARPS_PER_SM=32",   # shipped in round 1
    "s3_i8_c3": "-DRV_K1_STAGES=3 -DRV_K1_ITERS=8 -DRV_K1_WARPS_PER_SM=24",
    "s4_i8_c3": "-DRV_K1_STAGES=4 -DRV_K1_ITERS=8 -DRV_K1_WARPS_PER_SM=24",
    "s3_i6_c4": "-DRV_K1_STAGES=3 -DRV_K1_ITERS=6 -DRV_K1_WARPS_PER_SM=32",
    "s4_i4_c4": "-DRV_K1_STAGES=4 -DRV_K1_ITERS=4 -DRV_K1_WARPS_PER_SM=32",
    "s6_i4_c4": "-DRV_K1_STAGES=6 -DRV_K1_ITERS=4 -DRV_K1_WARPS_PER_SM=32",
    "s4_i4_c5": "-DRV_K1_STAGES=4 -DRV_K1_ITERS=4 -DRV_K1_WARPS_PER_SM=40",
    "s2_i8_c4_tb1": "-DRV_K1_STAGES=2 -DRV_K1_ITERS=8 -DRV_K1_WARPS_PER_SM=32 -DRV_K1_TICKET_BATCH=1",
    "s2_i8_c4_tb2": "-DRV_K1_STAGES=2 -DRV_K1_ITERS=8 -DRV_K1_WARPS_PER_SM=32 -DRV_K1_TICKET_BATCH=2",
    "s2_i8_c4_ahead": "-DRV_K1_STAGES=2 -DRV_K1_ITERS=8 -DRV_K1_WARPS_PER_SM=32 -DRV_K1_TICKET_AHEAD=1",
    # timing experiments without the prefix chain (offsets are wrong): what the chain costs at each ring depth
    "nochain_s2_i8_c4": "-DRV_K1_NO_CHAIN -DRV_K1_STAGES=2 -DRV_K1_ITERS=8 -DRV_K1_WARPS_PER_SM=32",
    "nochain_s4_i8_c3": "-DRV_K1_NO_CHAIN -DRV_K1_STAGES=4 -DRV_K1_ITERS=8 -DRV_K1_WARPS_PER_SM=24",
    "nochain_s6_i4_c4": "-DRV_K1_NO_CHAIN -DRV_K1_STAGES=6 -DRV_K1_ITERS=4 -DRV_K1_WARPS_PER_SM=32",
}


VARIANTS.update({"k1_opaque1": "-DRV_K1_OPAQUE=1", "k1_opaque0": "-DRV_K1_OPAQUE=0"})
# K4 tunables (tools/k4_sweep.sh): groups per warp step and occupancy of the insert, block size / occupancy of the emit
VARIANTS.update({
    "k4_g4_i3_e128x8": "-DRV_VOX_GROUPS=4 -DRV_VOX_INSERT_OCC=3 -DRV_VOX_EMIT_THREADS=128 -DRV_VOX_EMIT_OCC=8",
    "k4_g2_i5_e128x8": "-DRV_VOX_GROUPS=2 -DRV_VOX_INSERT_OCC=5 -DRV_VOX_EMIT_THREADS=128 -DRV_VOX_EMIT_OCC=8",
    "k4_g2_i6_e256x6": "-DRV_VOX_GROUPS=2 -DRV_VOX_INSERT_OCC=6 -DRV_VOX_EMIT_THREADS=256 -DRV_VOX_EMIT_OCC=6",
    "k4_g4_i4_e256x6": "-DRV_VOX_GROUPS=4 -DRV_VOX_INSERT_OCC=4 -DRV_VOX_EMIT_THREADS=256 -DRV_VOX_EMIT_OCC=6",
    "k4_g1_i6_e256x8": "-DRV_VOX_GROUPS=1 -DRV_VOX_INSERT_OCC=6 -DRV_VOX_EMIT_THREADS=256 -DRV_VOX_EMIT_OCC=8",
    "k4_g4_i3_e512x3": "-DRV_VOX_GROUPS=4 -DRV_VOX_INSERT_OCC=3 -DRV_VOX_EMIT_THREADS=512 -DRV_VOX_EMIT_OCC=3",
    "k4_g1_i8_e256x8": "-DRV_VOX_GROUPS=1 -DRV_VOX_INSERT_OCC=8 -DRV_VOX_EMIT_THREADS=256 -DRV_VOX_EMIT_OCC=8",
    "k4_g1_i6_e256x6": "-DRV_VOX_GROUPS=1 -DRV_VOX_INSERT_OCC=6 -DRV_VOX_EMIT_THREADS=256 -DRV_VOX_EMIT_OCC=6",
    "k4_g1_i6_e128x12": "-DRV_VOX_GROUPS=1 -DRV_VOX_INSERT_OCC=6 -DRV_VOX_EMIT_THREADS=128 -DRV_VOX_EMIT_OCC=12",
})
# more resident compute warps: 16 warps of 128 pixels on the 2048-pixel tile, or 1024-pixel tiles with more CTAs per SM
VARIANTS.update({
    "k1w_base": "",
    "k1w_cw16_i4_s2_w48": "-DRV_K1_CW=16 -DRV_K1_ITERS=4 -DRV_K1_STAGES=2 -DRV_K1_WARPS_PER_SM=48",
    "k1w_cw16_i4_s3_w48": "-DRV_K1_CW=16 -DRV_K1_ITERS=4 -DRV_K1_STAGES=3 -DRV_K1_WARPS_PER_SM=48",
    "k1w_cw8_i4_s2_w48": "-DRV_K1_CW=8 -DRV_K1_ITERS=4 -DRV_K1_STAGES=2 -DRV_K1_WARPS_PER_SM=48",
    "k1w_cw8_i4_s2_w56": "-DRV_K1_CW=8 -DRV_K1_ITERS=4 -DRV_K1_STAGES=2 -DRV_K1_WARPS_PER_SM=56",
    "k1w_cw8_i4_s3_w48": "-DRV_K1_CW=8 -DRV_K1_ITERS=4 -DRV_K1_STAGES=3 -DRV_K1_WARPS_PER_SM=48",
    "k1w_cw4_i8_s2_w40": "-DRV_K1_CW=4 -DRV_K1_ITERS=8 -DRV_K1_STAGES=2 -DRV_K1_WARPS_PER_SM=40",
})
VARIANTS.update({"k1p_ns0": "", "k1p_ns32": "-DRV_K1_POLL_NS=32", "k1p_ns100": "-DRV_K1_POLL_NS=100", "k1p_ns300": "-DRV_K1_POLL_NS=300"})
# ICP search tunables (tools/icp_sweep.sh): cell size of the nearest-point index in surface spacings, warm-started bound
VARIANTS.update({
    "icp_c10": "-DRV_NN_CELL_SPACINGS=1.0", "icp_c15": "-DRV_NN_CELL_SPACINGS=1.5", "icp_c20": "-DRV_NN_CELL_SPACINGS=2.0",
    "icp_c30": "-DRV_NN_CELL_SPACINGS=3.0", "icp_c40": "-DRV_NN_CELL_SPACINGS=4.0",
    "icp_c20_cold": "-DRV_NN_CELL_SPACINGS=2.0 -DRV_NN_WARM=0", "icp_c30_cold": "-DRV_NN_CELL_SPACINGS=3.0 -DRV_NN_WARM=0",
    "icp_c10_cold": "-DRV_NN_CELL_SPACINGS=1.0 -DRV_NN_WARM=0",
    "icp_c25": "-DRV_NN_CELL_SPACINGS=2.5", "icp_c35": "-DRV_NN_CELL_SPACINGS=3.5",
    "icp_occ9": "-DRV_NN_SEARCH_OCC=9", "icp_occ12": "-DRV_NN_SEARCH_OCC=12", "icp_occ14": "-DRV_NN_SEARCH_OCC=14",
    "icp_occ16": "-DRV_NN_SEARCH_OCC=16",
    "knn_f09": "-DRV_KNN_CELL_FACTOR=0.9", "knn_f12": "-DRV_KNN_CELL_FACTOR=1.2", "knn_f16": "-DRV_KNN_CELL_FACTOR=1.6",
    "knn_f20": "-DRV_KNN_CELL_FACTOR=2.0", "knn_f07": "-DRV_KNN_CELL_FACTOR=0.7",
    "knn_q1": "-DRV_KNN_QUERY_OCC=1", "knn_q10": "-DRV_KNN_QUERY_OCC=10", "knn_q12": "-DRV_KNN_QUERY_OCC=12",
    "icp_c30_timing": "-DRV_NN_CELL_SPACINGS=3.0 -DRV_ICP_TIMING",
})


def main():
    names = sys.argv[1:] or list(VARIANTS)
    outdir = os.path.join(ROOT, "build", "variants")
    os.makedirs(outdir, exist_ok=True)

    def one(n):
        return _build.build(out=os.path.join(outdir, f"librv_{n}.so"), extra=VARIANTS[n], tag="obj_" + n)

    with ThreadPoolExecutor(max_workers=3) as ex:
        for p in ex.map(one, names):
            print(p)


if __name__ == "__main__":
    main()
