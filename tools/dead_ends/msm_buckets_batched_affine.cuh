// Measured dead end of round 1, kept out of the product library (DESIGN.md "What bounds this path on B200"): batched-affine bucket
// accumulation.  Bit-exact on every parity test and 2.6x slower than the XYZZ kernel (profiles/r01_ncu_batched_affine_raw.csv.gz).
// To try it again: paste into csrc/msm.cu after k_msm_buckets and launch with
//   ZK_LAUNCH(k_msm_buckets_ba, ceil_div(M * K, ZK_BA_T * ZK_BA_K), ZK_BA_T, (size_t)ZK_BA_K * ZK_BA_T * 44, st, d_bases, ws.offsets.p,
//             ws.entries.p, D, M, ws.buckets.p, ws.order.p, ws.heavy_count.p, ws.heavy_list.p);
// (requires M * K and M * n * W below 2^32).
// Batched-affine bucket accumulation.  A mixed XYZZ addition costs 10 field products; an affine addition costs 3 plus one
// inversion, and Montgomery's trick shares one inversion among independent additions at 3 more products each.  Every thread
// owns ZK_BA_K buckets of (nearly) equal run length — consecutive ranks of the run-length order — and advances all of them one
// point per round: pass 1 multiplies up the denominators x_P - x_acc (prefix products parked in shared memory), one
// uniform-flow inversion (fe_inv), pass 2 walks back, peels off each denominator's inverse and finishes the additions:
// 6 products per addition + 1/ZK_BA_K of an inversion whose ALU-pipe work overlaps other warps' IMAD work.
// Accumulators live in the bucket array itself (x, y; L2-resident while a CTA works on them); ZZ = ZZZ = 1 is written at
// the end so k_msm_reduce reads ordinary XYZZ buckets.  The exceptional cases (identity base point, accumulator at infinity,
// P + P, P - P) leave the batch (denominator 1) and take a slow path with their own inversion, so the result is the exact
// group element in every case.
#define ZK_BA_K 16
#define ZK_BA_T 64
__global__ void __launch_bounds__(ZK_BA_T) k_msm_buckets_ba(const g1_affine_t* __restrict__ bases, const uint32_t* __restrict__ offsets,
                                                            const uint32_t* __restrict__ entries, MsmDims D, size_t M,
                                                            g1_xyzz_t* __restrict__ buckets, const uint32_t* __restrict__ order,
                                                            uint32_t* __restrict__ heavy_count, uint64_t* __restrict__ heavy_list) {
    extern __shared__ uint4 ba_sm[];
    const unsigned T = ZK_BA_T, tid = threadIdx.x;
    uint4* pre_lo = ba_sm;                    // prefix products, two 16-byte planes [K][T]
    uint4* pre_hi = ba_sm + ZK_BA_K * T;
    uint32_t* s_ent = reinterpret_cast<uint32_t*>(ba_sm + 2 * ZK_BA_K * T);   // first entry of the run (index into `entries`)
    uint32_t* s_len = s_ent + ZK_BA_K * T;                                     // run length (0: nothing to do)
    uint32_t* s_bkt = s_len + ZK_BA_K * T;                                     // bucket index
    const size_t K = (size_t)D.G * D.nb, total = M * K, per_msm = (size_t)D.n * D.W;
    const size_t base_rank = (size_t)blockIdx.x * (T * ZK_BA_K);
    const fq_t one = fe_one<FqTag>();
    uint32_t maxlen = 0, inf = 0, skip = 0;   // bit j of inf: accumulator j is the point at infinity; of skip: slot j is not ours
    for (unsigned j = 0; j < ZK_BA_K; ++j) {
        const size_t rank = base_rank + (size_t)j * T + tid;
        uint32_t len = 0, ent = 0, bkt = 0;
        if (rank >= total) skip |= 1u << j;
        else {
            const size_t m = rank / K;
            const uint32_t key = order[rank];
            const uint32_t* om = offsets + m * (K + 1);
            const uint32_t b = om[key], e = om[key + 1];
            bkt = (uint32_t)(m * K + key); ent = (uint32_t)(m * per_msm + b); len = e - b;
            if (len > D.heavy) { heavy_list[atomicAdd(heavy_count, 1u)] = (uint64_t)bkt; skip |= 1u << j; len = 0; }
            else if (len == 0) inf |= 1u << j;
            else {
                const uint32_t ref = entries[ent];
                const g1_affine_t* p = bases + (ref & 0x7fffffffu);
                fq_t px = fe_ldg(&p->x), py = fe_ldg(&p->y);
                if (px.is_zero() && py.is_zero()) inf |= 1u << j;
                else {
                    if (ref >> 31) py = neg(py);
                    fe_store(&buckets[bkt].x, px); fe_store(&buckets[bkt].y, py);
                }
            }
        }
        s_ent[j * T + tid] = ent; s_len[j * T + tid] = len; s_bkt[j * T + tid] = bkt;
        maxlen = len > maxlen ? len : maxlen;
    }
    for (uint32_t t = 1; t < maxlen; ++t) {
        fq_t prod = one;
        uint32_t special = 0;
#pragma unroll 2
        for (unsigned j = 0; j < ZK_BA_K; ++j) {
            if (t >= s_len[j * T + tid]) continue;
            const uint32_t ref = entries[s_ent[j * T + tid] + t];
            const fq_t px = fe_ldg(&bases[ref & 0x7fffffffu].x);
            if ((inf >> j) & 1u) { special |= 1u << j; continue; }
            const fq_t d = px - fe_load(&buckets[s_bkt[j * T + tid]].x);
            if (d.is_zero() || px.is_zero()) { special |= 1u << j; continue; }
            pre_lo[j * T + tid] = make_uint4(prod.l[0], prod.l[1], prod.l[2], prod.l[3]);
            pre_hi[j * T + tid] = make_uint4(prod.l[4], prod.l[5], prod.l[6], prod.l[7]);
            prod = prod * d;
        }
        fq_t inv = fe_inv(prod);
#pragma unroll 2
        for (unsigned jj = ZK_BA_K; jj-- > 0;) {
            const unsigned j = jj;
            if (t >= s_len[j * T + tid]) continue;
            const uint32_t ref = entries[s_ent[j * T + tid] + t];
            const g1_affine_t* p = bases + (ref & 0x7fffffffu);
            const fq_t px = fe_ldg(&p->x);
            fq_t py = fe_ldg(&p->y);
            g1_xyzz_t* acc = buckets + s_bkt[j * T + tid];
            if ((special >> j) & 1u) {
                // slow path: exact handling of the exceptional cases, own inversion
                if (px.is_zero() && py.is_zero()) continue;              // identity base point: nothing to add
                if (ref >> 31) py = neg(py);
                if ((inf >> j) & 1u) { fe_store(&acc->x, px); fe_store(&acc->y, py); inf &= ~(1u << j); continue; }
                const fq_t ax = fe_load(&acc->x), ay = fe_load(&acc->y);
                fq_t lam;
                if (px == ax) {
                    if (py == ay) { fq_t x2 = sqr(ax); lam = (dbl(x2) + x2) * fe_inv(dbl(ay)); }    // P + P
                    else { inf |= 1u << j; continue; }                                             // P - P
                } else lam = (py - ay) * fe_inv(px - ax);
                const fq_t x3 = sqr(lam) - ax - px;
                fe_store(&acc->y, lam * (ax - x3) - ay); fe_store(&acc->x, x3);
                continue;
            }
            if (ref >> 31) py = neg(py);
            const fq_t ax = fe_load(&acc->x), ay = fe_load(&acc->y);
            fq_t pre;
            { uint4 lo = pre_lo[j * T + tid], hi = pre_hi[j * T + tid];
              pre.l[0] = lo.x; pre.l[1] = lo.y; pre.l[2] = lo.z; pre.l[3] = lo.w; pre.l[4] = hi.x; pre.l[5] = hi.y; pre.l[6] = hi.z; pre.l[7] = hi.w; }
            const fq_t dinv = inv * pre;
            inv = inv * (px - ax);
            const fq_t lam = (py - ay) * dinv;
            const fq_t x3 = sqr(lam) - ax - px;
            fe_store(&acc->y, lam * (ax - x3) - ay); fe_store(&acc->x, x3);
        }
    }
    const fq_t zero = fq_t::zero();
    for (unsigned j = 0; j < ZK_BA_K; ++j) {
        if ((skip >> j) & 1u) continue;
        g1_xyzz_t* acc = buckets + s_bkt[j * T + tid];
        if ((inf >> j) & 1u) { fe_store(&acc->x, zero); fe_store(&acc->y, zero); fe_store(&acc->zz, zero); fe_store(&acc->zzz, zero); }
        else { fe_store(&acc->zz, one); fe_store(&acc->zzz, one); }
    }
}

