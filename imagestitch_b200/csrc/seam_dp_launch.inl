// seam_dp_launch.inl -- shapes and launches of the DP kernels (included by seam.cu inside namespace is, after the kernels):
// used by the batched path (seam_batch.inl), by the general per-pair path (PairSeam::estimate_and_update) and by the
// micro-benchmark entry.
// ---- DP launches ----------------------------------------------------------------------------------------------------------
// Two formulations of the forward pass (seam.cu): 0 = one __syncthreads per step over a TMA-fed shared-memory ring, back-track
// inside the kernel; 1 = warp-private halo windows, one __syncthreads per H steps, parallel back-track kernels.
struct DpShape {
    int v1 = 0;                        // formulation
    int tmpl = 0, nwarps = 0;          // v1: template choice, warps per CTA
    int cl = 1;                        // v1: CTAs per seam (thread-block cluster), 1 = a single CTA
    int lpt = 4, nt = 32;              // v0
    int pitch = 128, G = 1, D = 2;
    int lanes = 0, s0 = 0, s1 = 0;
    bool same(const DpShape& o) const { return v1 == o.v1 && (v1 ? (tmpl == o.tmpl && nwarps == o.nwarps && cl == o.cl) : (lpt == o.lpt && nt == o.nt)); }
};

static int dp_variant_default() {
    if (const char* e = getenv("IS_DP_VARIANT")) return atoi(e) ? 1 : 0;
    return 1;
}

// window shapes of k_seam_fwd<LPT, H, R>: {LPT, H, owned lanes per warp}
static const int V1_TMPL[4][3] = {{4, 8, 112}, {8, 16, 224}, {16, 16, 480}, {4, 16, 96}};

static void dp_choose_shape(int lanes, int steps, int s0, int s1, int variant, DpShape* S) {
    S->lanes = lanes; S->s0 = s0; S->s1 = s1;
    S->lpt = lanes <= 4096 ? 4 : (lanes <= 8192 ? 8 : 16);
    if (const char* e = getenv("IS_DP_LPT")) {
        const int v = atoi(e);
        if ((v == 4 || v == 8 || v == 16) && lanes <= 1024 * v) S->lpt = v;
    }
    S->nt = std::min(1024, div_up(div_up(lanes, S->lpt), 32) * 32);
    S->pitch = S->nt * S->lpt;
    const size_t row_pair = 2 * sizeof(float) * (size_t)S->pitch;
    S->D = 2;
    S->G = (int)std::min<size_t>(16, (192 * 1024) / (S->D * row_pair));
    if (S->G < 1) S->G = 1;
    if (const char* e = getenv("IS_DP_G")) S->G = std::max(1, atoi(e));
    if (const char* e = getenv("IS_DP_D")) S->D = std::max(2, atoi(e));
    S->v1 = 0;
    if (variant == 1) {
        int forced = -1;
        if (const char* e = getenv("IS_DP_V1_TMPL")) forced = atoi(e);
        static const int order[4] = {3, 0, 1, 2};          // the 16-step halo first (half as many block barriers), then by capacity
        for (int k = 0; k < 4; ++k) {
            const int t = order[k];
            if (forced >= 0 && t != forced) continue;
            if (lanes <= 16 * V1_TMPL[t][2]) { S->v1 = 1; S->tmpl = t; S->nwarps = div_up(lanes, V1_TMPL[t][2]); break; }
        }
        // a seam of more than four windows is spread over a cluster of 2, 4 or 8 CTAs (k_seam_fwd_cluster, window shape {4, 16, 96}):
        // as few windows per CTA as the cluster size allows -- one per scheduler up to 3072 lanes
        int cl = 0;
        if (const char* e = getenv("IS_DP_CLUSTER")) cl = atoi(e);                  // tuning knob: 1 = never, 2 / 4 / 8 = always that size
        const int win = div_up(lanes, V1_TMPL[3][2]);
        if (forced < 0 && cl != 1 && win <= 8 * 16 && (cl == 2 || cl == 4 || cl == 8 || win > 4)) {
            if (cl != 2 && cl != 4 && cl != 8) cl = win <= 8 ? 2 : (win <= 16 ? 4 : 8);     // four windows per CTA (ring of 16 steps) up to 32 windows
            if (div_up(win, cl) <= 16) { S->v1 = 1; S->tmpl = 3; S->cl = cl; S->nwarps = div_up(win, cl); }
        }
    }
    (void)steps;
}

template <int LPT>
static int launch_dp_v0(is_ctx* ctx, const DpArgs* table_d, int njobs, int nt, size_t smem, double bytes) {
    IS_CUDA(ctx, cudaFuncSetAttribute(k_seam_dp_batch<LPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, DP_SMEM_MAX));
    ctx->next_bytes = bytes;
    IS_LAUNCH(ctx, k_seam_dp_batch<LPT>, njobs, nt, smem, table_d);
    return IS_OK;
}

// shapes[q] describes table entry q; entries of equal shape are adjacent.  bt_d: BtArgs of the v1 entries (same indices).
static int launch_dp_all(is_ctx* ctx, const std::vector<DpShape>& shapes, const DpArgs* dp_d, const BtArgs* bt_d) {
    const size_t nj = shapes.size();
    int v1_first = -1, v1_count = 0, max_lanes = 0, max_chunks = 0;
    for (size_t q = 0; q < nj;) {
        size_t e = q;
        const DpShape& S0 = shapes[q];
        double bytes = 0;
        while (e < nj && shapes[e].same(S0)) { bytes += (double)(shapes[e].s1 - shapes[e].s0) * shapes[e].lanes * 9; ++e; }
        const int cnt = (int)(e - q);
        if (!S0.v1) {
            const size_t row_pair = 2 * sizeof(float) * (size_t)S0.pitch;
            const size_t smem = std::max<size_t>((size_t)S0.D * S0.G * row_pair + 8 * (size_t)S0.D + 16, (size_t)32 * 65 + 16);
            IS_REQUIRE(ctx, smem <= (size_t)DP_SMEM_MAX, IS_ERR_INTERNAL, "DP shared-memory budget");
            switch (S0.lpt) {
                case 4: IS_TRY(launch_dp_v0<4>(ctx, dp_d + q, cnt, S0.nt, smem, bytes)); break;
                case 8: IS_TRY(launch_dp_v0<8>(ctx, dp_d + q, cnt, S0.nt, smem, bytes)); break;
                default: IS_TRY(launch_dp_v0<16>(ctx, dp_d + q, cnt, S0.nt, smem, bytes)); break;
            }
        } else {
            const int H = V1_TMPL[S0.tmpl][1];
            const size_t smem = 2 * sizeof(float) * (size_t)(S0.pitch + 2 * H);
            ctx->next_bytes = bytes;
            if (S0.cl > 1) {
                const size_t csmem = 2 * sizeof(float) * (size_t)(S0.nwarps * V1_TMPL[3][2] + 2 * H);
                int ring = S0.nwarps <= 4 ? 16 : (S0.nwarps <= 8 ? 8 : 4);             // steps of cost rows in registers ahead of their use
                if (const char* e = getenv("IS_DP_CL_RING")) ring = std::min(ring, atoi(e));   // tuning knob: a shallower ring (4, 8)
                const bool asy = getenv("IS_DP_CL_BARRIER") == nullptr;                   // tuning knob: halo exchange behind cluster barriers instead of st.async
#define IS_DP_CL_LAUNCH(CLN, RN, MT)                                                                                                             \
    do {                                                                                                                                         \
        if (asy) IS_LAUNCH(ctx, (k_seam_fwd_cluster<4, 16, RN, CLN, MT, true>), cnt * CLN, S0.nwarps * 32, csmem, dp_d + q);                     \
        else IS_LAUNCH(ctx, (k_seam_fwd_cluster<4, 16, RN, CLN, MT, false>), cnt * CLN, S0.nwarps * 32, csmem, dp_d + q);                        \
    } while (0)
#define IS_DP_CL_CASE(CLN)                                                                                                                       \
    case CLN:                                                                                                                                    \
        if (ring == 16) IS_DP_CL_LAUNCH(CLN, 16, 128);                                                                                           \
        else if (ring == 8) IS_DP_CL_LAUNCH(CLN, 8, 256);                                                                                        \
        else IS_DP_CL_LAUNCH(CLN, 4, 512);                                                                                                       \
        break;
                switch (S0.cl) {
                    IS_DP_CL_CASE(2)
                    IS_DP_CL_CASE(4)
                    default:
                    IS_DP_CL_CASE(8)
                }
#undef IS_DP_CL_CASE
#undef IS_DP_CL_LAUNCH
            } else if (S0.tmpl == 3 && S0.nwarps <= 4 && !getenv("IS_DP_CL_RING")) {             // a single CTA of at most four windows: the deep ring as well
                IS_LAUNCH(ctx, (k_seam_fwd<4, 16, 16, 128>), cnt, S0.nwarps * 32, smem, dp_d + q);
            } else
            switch (S0.tmpl) {
                case 0: IS_LAUNCH(ctx, (k_seam_fwd<4, 8, 4>), cnt, S0.nwarps * 32, smem, dp_d + q); break;
                case 1: IS_LAUNCH(ctx, (k_seam_fwd<8, 16, 2>), cnt, S0.nwarps * 32, smem, dp_d + q); break;
                case 2: IS_LAUNCH(ctx, (k_seam_fwd<16, 16, 1>), cnt, S0.nwarps * 32, smem, dp_d + q); break;
                default: IS_LAUNCH(ctx, (k_seam_fwd<4, 16, 4>), cnt, S0.nwarps * 32, smem, dp_d + q); break;
            }
            if (v1_first < 0) v1_first = (int)q;
            v1_count = (int)e - v1_first;                               // v1 entries are contiguous (sorted by formulation first)
            for (size_t k = q; k < e; ++k) { max_lanes = std::max(max_lanes, shapes[k].lanes); max_chunks = std::max(max_chunks, div_up(shapes[k].s1 - shapes[k].s0, BT_CHUNK)); }
        }
        q = e;
    }
    if (v1_count > 0 && max_chunks > 0) {
        IS_LAUNCH(ctx, k_bt_compose, dim3(div_up(max_lanes, 256), max_chunks, v1_count), 256, 0, bt_d + v1_first);
        IS_LAUNCH(ctx, k_bt_walk, v1_count, 256, sizeof(int) * (size_t)max_chunks, bt_d + v1_first);
    }
    return IS_OK;
}

