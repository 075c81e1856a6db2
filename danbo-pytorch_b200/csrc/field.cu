// Near/far (NF1, NF2), coarse sampling (SM1), world->bone transform + visibility mask (T1, T2, G1 mask),
// per-bone feature gather + aggregation net + blend + positional encoding (G1, G2, A1-A3), per-ray view bias (V1).
// Reference rows: SURVEY.md §8(a).  All kernels are HBM / latency bound integer-and-fp32 work; they share one
// layout rule: one thread (or one warp) per ray sample, pose tables read through L1 as warp-uniform float4 loads.
#include "field_common.cuh"

namespace danbo {

// ---------------------------------------------------------------------------------------------------------
// NF1: ray / bounding-cylinder near & far in the x-z plane, fp32, op order of ray_utils.py:294-328.
// Writes NaN for rays that miss; accumulates per-segment sums for the reference's chunk-wide nanmean (F8).
__global__ void nearfar_cyl_kernel(const float* __restrict__ rays, int ray_stride, int n_rays,
                                   const float* __restrict__ pose_cyl, int cyl_stride, int rays_per_pose, int n_poses,
                                   int seg_len, float* __restrict__ near_out, float* __restrict__ far_out,
                                   double* __restrict__ seg_acc /* [n_seg][4] sum_near,cnt_near,sum_far,cnt_far */) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_rays) return;                             // (tail warps fall back to per-thread atomics below)
    const float* r = rays + (size_t)n * ray_stride;
    const float ox = r[0], oz = r[2], dx = r[3], dz = r[5], near = r[6], far = r[7];
    int pose = n / rays_per_pose; if (pose >= n_poses) pose = n_poses - 1;
    const float* c = pose_cyl + (size_t)pose * cyl_stride;
    const float cx = c[0], cz = c[1], rad = c[2];
    const float nx = __fadd_rn(ox, __fmul_rn(dx, near)), nz = __fadd_rn(oz, __fmul_rn(dz, near));
    const float fx = __fadd_rn(ox, __fmul_rn(dx, far)), fz = __fadd_rn(oz, __fmul_rn(dz, far));
    const float ncx = __fsub_rn(cx, nx), ncz = __fsub_rn(cz, nz);
    const float nfx = __fsub_rn(fx, nx), nfz = __fsub_rn(fz, nz);
    const float nf_norm = __fsqrt_rn(__fadd_rn(__fmul_rn(nfx, nfx), __fmul_rn(nfz, nfz)));
    const float scale = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dz, dz)));
    const float cross = __fsub_rn(__fmul_rn(ncx, nfz), __fmul_rn(ncz, nfx));
    const float dist = __fdiv_rn(fabsf(cross), nf_norm);
    const float Q = __fsqrt_rn(__fsub_rn(__fmul_rn(rad, rad), __fmul_rn(dist, dist)));      // NaN when the ray misses
    const float K = __fdiv_rn(__fadd_rn(__fmul_rn(ncx, nfx), __fmul_rn(ncz, nfz)), nf_norm);
    const float m = (Q < K) ? 1.f : 0.f;
    const float nn = __fadd_rn(near, __fdiv_rn(__fmul_rn(m, __fsub_rn(K, Q)), scale));
    const float ff = __fadd_rn(near, __fdiv_rn(__fadd_rn(K, Q), scale));
    // rows with NaN Q are refilled later; flag them by writing NaN to BOTH outputs' sign of Q via far (already NaN)
    near_out[n] = (Q != Q) ? __int_as_float(0x7fc00000) : nn;
    far_out[n] = (Q != Q) ? __int_as_float(0x7fc00000) : ff;
    const int seg = seg_len > 0 ? n / seg_len : 0;
    double* acc = seg_acc + 4 * (size_t)seg;
    // nanmean over the segment: every non-NaN entry counts (ray_utils.py:334,340).  One atomic per warp and quantity
    // when the warp sits inside one segment (always, for the reference's 4096-ray chunks).
    double v0 = (nn == nn) ? (double)nn : 0.0, c0 = (nn == nn) ? 1.0 : 0.0;
    double v1 = (ff == ff) ? (double)ff : 0.0, c1 = (ff == ff) ? 1.0 : 0.0;
    const unsigned active = __activemask();
    const int seg0 = __shfl_sync(active, seg, __ffs(active) - 1);
    if (__all_sync(active, seg == seg0) && active == 0xffffffffu) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v0 += __shfl_xor_sync(0xffffffffu, v0, o); c0 += __shfl_xor_sync(0xffffffffu, c0, o);
            v1 += __shfl_xor_sync(0xffffffffu, v1, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (c0 > 0.0) { atomicAdd(acc + 0, v0); atomicAdd(acc + 1, c0); }
            if (c1 > 0.0) { atomicAdd(acc + 2, v1); atomicAdd(acc + 3, c1); }
        }
    } else {
        if (c0 > 0.0) { atomicAdd(acc + 0, v0); atomicAdd(acc + 1, c0); }
        if (c1 > 0.0) { atomicAdd(acc + 2, v1); atomicAdd(acc + 3, c1); }
    }
}

// NF1 fill + NF2: per-bone oriented-box near/far in fp64 with the exactly-two-hits rule (F7).
// raycasters.py:648-707, ray_utils.py:383-417.  One thread per ray.
__global__ void nearfar_finish_kernel(const float* __restrict__ rays, int ray_stride, int n_rays,
                                      const float* __restrict__ pose_skts, int rays_per_pose, int n_poses,
                                      FieldConsts fc, int seg_len, const double* __restrict__ seg_acc, int use_box,
                                      float bound, float hi, float* __restrict__ near_io, float* __restrict__ far_io,
                                      uint8_t* __restrict__ p_valid /* (N,24) 6-bit masks or null */,
                                      uint8_t* __restrict__ v_valid /* (N,24) or null */) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_rays) return;
    const float* r = rays + (size_t)n * ray_stride;
    float near = near_io[n], far = far_io[n];
    if (far != far) {                                   // Q was NaN: take the segment nanmean, else the input bound
        const double* acc = seg_acc + 4 * (size_t)(seg_len > 0 ? n / seg_len : 0);
        near = acc[1] > 0.0 ? (float)(acc[0] / acc[1]) : r[6];
        far = acc[3] > 0.0 ? (float)(acc[2] / acc[3]) : r[7];
    }
    if (use_box) {
        int pose = n / rays_per_pose; if (pose >= n_poses) pose = n_poses - 1;
        const float ox = r[0], oy = r[1], oz = r[2], dx = r[3], dy = r[4], dz = r[5];
        // hi = fp32(bound_range + eps) computed by the host in double, the way torch compares a float tensor
        // with the Python scalar `bound_range + eps` (ray_utils.py:404-409)
        float vnear = 100000.f, vfar = -100000.f;
        bool any = false;
        for (int j = 0; j < DANBO_J; ++j) {
            const float* S = pose_skts + ((size_t)pose * DANBO_J + j) * 16;
            const float* A = fc.align + j * 16;
            // rays_ot = R o + t ; rays_dt = R d            (batched 3x3 matvec, fp32)
            float o[3], d[3], ot[3], dt[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                o[i] = __fadd_rn(fmaf(S[4 * i + 2], oz, fmaf(S[4 * i + 1], oy, __fmul_rn(S[4 * i], ox))), S[4 * i + 3]);
                d[i] = fmaf(S[4 * i + 2], dz, fmaf(S[4 * i + 1], dy, __fmul_rn(S[4 * i], dx)));
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                ot[i] = __fadd_rn(fmaf(A[4 * i + 2], o[2], fmaf(A[4 * i + 1], o[1], __fmul_rn(A[4 * i], o[0]))), A[4 * i + 3]);
                dt[i] = fmaf(A[4 * i + 2], d[2], fmaf(A[4 * i + 1], d[1], __fmul_rn(A[4 * i], d[0])));
            }
            float s[3], os[3], ds[3];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                s[i] = fabsf(fc.axis_scale[j * 3 + i]);
                os[i] = __fdiv_rn(ot[i], s[i]);
                ds[i] = __fdiv_rn(dt[i], s[i]);
            }
            int n_in = 0;
            uint32_t bits = 0;
            float seg[2][3];
            // fp32 slab test of the whole LINE against the box grown by a margin far above fp32 rounding: if even that
            // misses, none of the six plane points below can lie inside the box (n_in = 0, the bone is not valid), so the
            // fp64 plane intersections are skipped - a ray of the 512x512 image comes near 3-5 of the 24 boxes.
            bool may_hit = true;
            {
                const float H = hi * 1.002f + 2e-3f;
                float lo_t = -3.0e38f, hi_t = 3.0e38f;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    if (fabsf(ds[i]) < 1e-12f) { may_hit = may_hit && !(fabsf(os[i]) > H); continue; }
                    const float inv = 1.f / ds[i];
                    const float ta = (-H - os[i]) * inv, tb = (H - os[i]) * inv;
                    lo_t = fmaxf(lo_t, fminf(ta, tb)); hi_t = fminf(hi_t, fmaxf(ta, tb));
                }
                // widen the interval test itself: t values are O(1..10), their fp32 error O(1e-6)
                may_hit = may_hit && !(lo_t > hi_t + 1e-3f * (1.f + fabsf(lo_t) + fabsf(hi_t)));
            }
            if (may_hit)
#pragma unroll
            for (int k = 0; k < 6; ++k) {               // planes: -b on x,y,z then +b on x,y,z
                const int a = k % 3;
                const double b = (k < 3) ? -(double)bound : (double)bound;
                const double t = (b - (double)os[a]) / (double)ds[a];
                float p[3];
                bool in = true;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    p[i] = (float)(t * (double)ds[i] + (double)os[i]);
                    in = in && (p[i] <= hi) && (p[i] >= -hi);
                }
                if (in) {
                    if (n_in < 2) { seg[n_in][0] = p[0]; seg[n_in][1] = p[1]; seg[n_in][2] = p[2]; }
                    ++n_in;
                    bits |= 1u << k;
                }
            }
            const bool valid = (n_in == 2);
            if (p_valid) p_valid[(size_t)n * DANBO_J + j] = (uint8_t)bits;
            if (v_valid) v_valid[(size_t)n * DANBO_J + j] = valid ? 1 : 0;
            if (valid) {
                const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dt[0], dt[0]), __fmul_rn(dt[1], dt[1])), __fmul_rn(dt[2], dt[2])));
                float st[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float q0 = __fsub_rn(__fmul_rn(seg[e][0], s[0]), ot[0]);
                    const float q1 = __fsub_rn(__fmul_rn(seg[e][1], s[1]), ot[1]);
                    const float q2 = __fsub_rn(__fmul_rn(seg[e][2], s[2]), ot[2]);
                    st[e] = __fdiv_rn(__fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)), __fmul_rn(q2, q2))), nrm);
                }
                vnear = fminf(vnear, fminf(st[0], st[1]));
                vfar = fmaxf(vfar, fmaxf(st[0], st[1]));
                any = true;
            }
        }
        if (any) { near = vnear; far = vfar; }
    }
    near_io[n] = near;
    far_io[n] = far;
}

// ---------------------------------------------------------------------------------------------------------
// SM1 + T1/T2 + visibility mask + compaction.  One thread per (ray, sample).
//   coarse mode (z_in == null): z = near(1-t)+far*t (+ stratified jitter), written to z_out
//   fine mode   (z_in != null): z given (importance samples)
// Emits the 24-bit visibility mask of every sample and appends samples with a non-empty mask to the active
// list; with append_empty, one extra entry per ray (id = n_rays*S + ray) stands for "a sample no bone sees".
// kTable: the block first builds, per ray it covers and per bone, the affine map z -> x = c + z e of the two-step
// transform (x is affine in the sample depth), then every sample needs 3 FMAs per bone instead of ~45 instructions.
// Rounding differs from the reference's op order by ~1e-5 at most, so a sample whose largest |coordinate| lies within
// 2e-4 of a box face is re-evaluated in the exact order: the emitted mask is identical to the direct evaluation.
constexpr int kMaskBlock = 256;
constexpr int kMaxRaysPerBlock = 18;       // 256 / 16 + 2

// Sets bit j of *cand unless the affine map x(z) = cc + z ee stays outside |x|_inf <= 1 + kCull for every depth z.
__device__ __forceinline__ void mark_candidate(uint32_t* cand, int j, const float (&cc)[3], const float (&ee)[3]) {
    constexpr float kCull = 4e-4f;                                           // > the 2e-4 decision margin + interval rounding
    float lo = -3.0e38f, hi = 3.0e38f;
    bool empty = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        if (fabsf(ee[i]) < 1e-20f) { empty |= fabsf(cc[i]) > 1.f + kCull; continue; }
        const float inv = 1.f / ee[i];
        const float a = (-(1.f + kCull) - cc[i]) * inv, b = ((1.f + kCull) - cc[i]) * inv;
        lo = fmaxf(lo, fminf(a, b)); hi = fminf(hi, fmaxf(a, b));
    }
    if (!(empty || lo > hi)) atomicOr(cand, 1u << j);
}

template <bool kTable>
__global__ void __launch_bounds__(kMaskBlock)
sample_mask_kernel(const float* __restrict__ rays, int ray_stride, int n_rays, int S,
                                   const float* __restrict__ near, const float* __restrict__ far,
                                   const float* __restrict__ t_vals, const float* __restrict__ t_rand,
                                   const float* __restrict__ z_in, float* __restrict__ z_out,
                                   const float* __restrict__ pose_skts, int rays_per_pose, int n_poses,
                                   FieldConsts fc, uint32_t* __restrict__ mask_out, int* __restrict__ active_ids,
                                   int* __restrict__ active_count, int capacity, int append_empty, int lindisp) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = n_rays * S;
    uint32_t mask = 0;
    int n = 0, s = 0;
    __shared__ float4 tab[kTable ? kMaxRaysPerBlock * DANBO_J * 2 : 1];     // (c.xyz, e.xyz) per (ray slot, bone)
    const int ray_first = (blockIdx.x * kMaskBlock) / S;
    __shared__ float4 pm[kTable ? DANBO_J * 3 : 1];                          // per bone: rows of [diag(1/|s|) A R | offset]
    // cand[slot] bit j: the ray of this slot comes within kCull of bone j's box at SOME depth.  For every other bone the
    // table value max|x| exceeds 1 + kCull at every z, i.e. "decided, outside" below: those bones are never visited
    // (a ray of the 512x512 image comes near 3-5 of the 24 boxes, and half of the rays near none).
    __shared__ uint32_t cand[kTable ? kMaxRaysPerBlock : 1];
    if (kTable) {
        if (threadIdx.x < kMaxRaysPerBlock) cand[threadIdx.x] = 0u;
        const int ray_last = min(n_rays - 1, (blockIdx.x * kMaskBlock + kMaskBlock - 1) / S);
        const int n_ent = (ray_last - ray_first + 1) * DANBO_J;
        int pose_a = ray_first / rays_per_pose; if (pose_a >= n_poses) pose_a = n_poses - 1;
        int pose_b = ray_last / rays_per_pose; if (pose_b >= n_poses) pose_b = n_poses - 1;
        if (pose_a == pose_b) {
            // one pose in this block (always in a render call): fold the two affine steps and the scale into one 3x4
            // map per bone first (24 threads), then an entry costs two 3x3 products against shared memory.  The table
            // is approximate by construction (faces within 2e-4 are re-evaluated exactly below), so folding is free.
            if (threadIdx.x < DANBO_J) {
                const int j = threadIdx.x;
                const float4* sk4 = reinterpret_cast<const float4*>(pose_skts + ((size_t)pose_a * DANBO_J + j) * 16);
                const float4* A4 = reinterpret_cast<const float4*>(fc.align + j * 16);
                const float4 r0 = __ldg(sk4), r1 = __ldg(sk4 + 1), r2 = __ldg(sk4 + 2);
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float inv = 1.f / fabsf(__ldg(fc.axis_scale + j * 3 + i));
                    const float4 a = __ldg(A4 + i);
                    pm[j * 3 + i] = make_float4((a.x * r0.x + a.y * r1.x + a.z * r2.x) * inv, (a.x * r0.y + a.y * r1.y + a.z * r2.y) * inv,
                                                (a.x * r0.z + a.y * r1.z + a.z * r2.z) * inv,
                                                (a.x * r0.w + a.y * r1.w + a.z * r2.w + a.w) * inv);
                }
            }
            __syncthreads();
            for (int e = threadIdx.x; e < n_ent; e += kMaskBlock) {
                const int slot = e / DANBO_J, j = e - slot * DANBO_J;
                const float* r = rays + (size_t)(ray_first + slot) * ray_stride;
                const float o0 = r[0], o1 = r[1], o2 = r[2], d0 = r[3], d1 = r[4], d2 = r[5];
                const float4 m0 = pm[j * 3], m1 = pm[j * 3 + 1], m2 = pm[j * 3 + 2];
                const float cc[3] = {m0.x * o0 + m0.y * o1 + m0.z * o2 + m0.w, m1.x * o0 + m1.y * o1 + m1.z * o2 + m1.w,
                                     m2.x * o0 + m2.y * o1 + m2.z * o2 + m2.w};
                const float ee[3] = {m0.x * d0 + m0.y * d1 + m0.z * d2, m1.x * d0 + m1.y * d1 + m1.z * d2,
                                     m2.x * d0 + m2.y * d1 + m2.z * d2};
                tab[2 * e] = make_float4(cc[0], cc[1], cc[2], 0.f);
                tab[2 * e + 1] = make_float4(ee[0], ee[1], ee[2], 0.f);
                mark_candidate(&cand[slot], j, cc, ee);
            }
        } else {
            __syncthreads();                                              // cand[] zeroed
            for (int e = threadIdx.x; e < n_ent; e += kMaskBlock) {
                const int slot = e / DANBO_J, j = e - slot * DANBO_J;
                const int rn = ray_first + slot;
                const float* r = rays + (size_t)rn * ray_stride;
                int pose = rn / rays_per_pose; if (pose >= n_poses) pose = n_poses - 1;
                // 16-byte loads: lanes walk over bones (64 B apart), so every load instruction costs one LSU pass per
                // touched line whatever its width -- scalar loads made this table build the kernel's bottleneck
                const float4* sk4 = reinterpret_cast<const float4*>(pose_skts + ((size_t)pose * DANBO_J + j) * 16);
                const float4* A4 = reinterpret_cast<const float4*>(fc.align + j * 16);
                const float o0 = r[0], o1 = r[1], o2 = r[2], d0 = r[3], d1 = r[4], d2 = r[5];
                float c[3], d[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float4 row = __ldg(sk4 + i);
                    c[i] = row.x * o0 + row.y * o1 + row.z * o2 + row.w;
                    d[i] = row.x * d0 + row.y * d1 + row.z * d2;
                }
                float cc[3], ee[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float inv = 1.f / fabsf(__ldg(fc.axis_scale + j * 3 + i));
                    const float4 a = __ldg(A4 + i);
                    cc[i] = (a.x * c[0] + a.y * c[1] + a.z * c[2] + a.w) * inv;
                    ee[i] = (a.x * d[0] + a.y * d[1] + a.z * d[2]) * inv;
                }
                tab[2 * e] = make_float4(cc[0], cc[1], cc[2], 0.f);
                tab[2 * e + 1] = make_float4(ee[0], ee[1], ee[2], 0.f);
                mark_candidate(&cand[slot], j, cc, ee);
            }
        }
        __syncthreads();
    }
    if (idx < total) {
        n = idx / S; s = idx - n * S;
        const float* r = rays + (size_t)n * ray_stride;
        float z;
        if (z_in) {
            z = z_in[idx];
        } else {
            // lindisp (ray_utils.py:226-227): linear in inverse depth, 1/(1/near (1-t) + 1/far t); torch's `1./x` is a
            // correctly rounded reciprocal, so __frcp_rn reproduces it bit for bit
            const float nr = lindisp ? __frcp_rn(near[n]) : near[n], fr = lindisp ? __frcp_rn(far[n]) : far[n];
            auto zv = [&](int i) {
                const float t = t_vals[i];
                const float v = __fadd_rn(__fmul_rn(nr, __fsub_rn(1.f, t)), __fmul_rn(fr, t));
                return lindisp ? __frcp_rn(v) : v;
            };
            z = zv(s);
            if (t_rand) {                                // ray_utils.py:233-248
                const float lower = (s == 0) ? z : __fmul_rn(.5f, __fadd_rn(z, zv(s - 1)));
                const float upper = (s == S - 1) ? z : __fmul_rn(.5f, __fadd_rn(zv(s + 1), z));
                z = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), t_rand[idx]));
            }
            z_out[idx] = z;
        }
        const float px = __fadd_rn(r[0], __fmul_rn(r[3], z));
        const float py = __fadd_rn(r[1], __fmul_rn(r[4], z));
        const float pz = __fadd_rn(r[2], __fmul_rn(r[5], z));
        int pose = n / rays_per_pose; if (pose >= n_poses) pose = n_poses - 1;
        const float* skt = pose_skts + (size_t)pose * DANBO_J * 16;
        const float4* tr = tab + (kTable ? (n - ray_first) * DANBO_J * 2 : 0);
        for (uint32_t todo = kTable ? cand[n - ray_first] : (1u << DANBO_J) - 1u; todo; todo &= todo - 1) {
            const int j = __ffs(todo) - 1;
            bool decided = false, invalid = false;
            if (kTable) {
                // the largest |coordinate| decides: clearly above 1 -> outside (whatever the other two are), clearly below
                // -> inside; only a maximum within 2e-4 of the face needs the exact evaluation
                const float4 c = tr[2 * j], e = tr[2 * j + 1];
                const float am = fmaxf(fmaxf(fabsf(fmaf(z, e.x, c.x)), fabsf(fmaf(z, e.y, c.y))), fabsf(fmaf(z, e.z, c.z)));
                invalid = am > 1.f;
                decided = fabsf(am - 1.f) > 2e-4f;
            }
            if (!decided) {
                // |fl(t/s)| > 1  <=>  |t| > |s| for correctly rounded division (t > s implies t/s > 1 + 2^-24, which
                // rounds above 1), so the visibility mask needs no divide; field_agg divides for the few bones it reads
                float t0, t1, t2;
                bone_aligned(skt + j * 16, fc.align + j * 16, px, py, pz, t0, t1, t2);
                const float* sc = fc.axis_scale + j * 3;
                invalid = (fabsf(t0) > fabsf(__ldg(sc))) || (fabsf(t1) > fabsf(__ldg(sc + 1))) || (fabsf(t2) > fabsf(__ldg(sc + 2)));
            }
            mask |= (invalid ? 0u : 1u) << j;
        }
        mask_out[idx] = mask;
    }
    // ---- compaction: active samples, plus one "empty" entry per ray
    const bool act = (idx < total) && (mask != 0);
    // append_empty: 1 = one empty entry per ray, 2 = a single one (for the last ray; density queries)
    const bool emp = (idx < total) && (s == 0) && (append_empty == 1 || (append_empty == 2 && n == n_rays - 1));
    const int want = (act ? 1 : 0) + (emp ? 1 : 0);
    __shared__ int warp_tot[32];
    __shared__ int block_base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = want;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        int v = lane < nw ? warp_tot[lane] : 0, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
        if (lane < nw) warp_tot[lane] = inc - v;        // exclusive offset of each warp
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        if (lane == 0) block_base = tot > 0 ? atomicAdd(active_count, tot) : 0;
    }
    __syncthreads();
    if (want) {
        int pos = block_base + warp_tot[warp] + incl - want;
        if (act) { if (pos < capacity) active_ids[pos] = idx; ++pos; }
        if (emp && pos < capacity) active_ids[pos] = total + n;
    }
}

// ---------------------------------------------------------------------------------------------------------
// G1/G2 + A1-A3 + PE in three steps:
//   pair_count / pair_scatter : bucket the (row, visible bone) pairs by bone, segments padded to 32 pairs
//   pair_logits               : lane = one pair, warp = 32 pairs of the SAME bone, so every weight of the aggregation
//                               net is a warp-uniform (broadcast) load and each instruction does 32 useful MACs
//   field_rows                : lane = one row: blend weights, blended feature, positional encoding, bf16 X row
// An earlier version with lanes = hidden units spent 1 200 warp instructions per row on shuffles and weight loads
// (ncu: profiles/r1_ncu_field_summary.txt); this layout needs ~150.
template <bool kScatter>
__global__ void __launch_bounds__(256)
pair_bucket_kernel(const uint32_t* __restrict__ mask, const int* __restrict__ active_ids,
                   const int* __restrict__ active_count, int capacity, int total, PairWork pw, int pair_capacity,
                   int dense) {
    int count = *active_count; if (count > capacity) count = capacity;
    const int lane = threadIdx.x & 31;
    __shared__ int hist[DANBO_J];        // pairs of this block's current batch, per bone
    __shared__ int base[DANBO_J];        // scatter: where this block's pairs of bone j start
    for (int b0 = blockIdx.x * blockDim.x; b0 < count; b0 += gridDim.x * blockDim.x) {
        if (threadIdx.x < DANBO_J) hist[threadIdx.x] = 0;
        __syncthreads();
        const int e = b0 + threadIdx.x;
        uint32_t m = 0;
        if (e < count) { const int id = active_ids[e]; if (id < total) m = mask[id]; }
        // dense (softmax aggregation): the max over the logits runs over all 24 bones (danbo.py:399), so a row seen
        // by any bone needs the logit of every bone
        if (dense && m) m = (1u << DANBO_J) - 1u;
        uint32_t any = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, o);
        // pass 1: per-warp counts -> block histogram; remember this warp's offset inside the block's range
        int my_off = 0;                   // lane j < 24 keeps the warp's offset for bone j
        for (uint32_t a = any; a;) {
            const int j = __ffs(a) - 1; a &= a - 1;
            const uint32_t b = __ballot_sync(0xffffffffu, (m >> j) & 1u);
            int off = 0;
            if (lane == 0) off = atomicAdd(&hist[j], __popc(b));
            off = __shfl_sync(0xffffffffu, off, 0);
            if (lane == j) my_off = off;
        }
        __syncthreads();
        if (threadIdx.x < DANBO_J && hist[threadIdx.x]) {
            if (!kScatter) atomicAdd(pw.count() + threadIdx.x, hist[threadIdx.x]);
            else base[threadIdx.x] = seg_start(pw.count(), threadIdx.x) + atomicAdd(pw.cursor() + threadIdx.x, hist[threadIdx.x]);
        }
        if (kScatter) {
            __syncthreads();
            for (uint32_t a = any; a;) {
                const int j = __ffs(a) - 1; a &= a - 1;
                const uint32_t b = __ballot_sync(0xffffffffu, (m >> j) & 1u);
                const int woff = __shfl_sync(0xffffffffu, my_off, j);
                if ((m >> j) & 1u) {
                    const int at = base[j] + woff + __popc(b & ((1u << lane) - 1));
                    if (at < pair_capacity) pw.pairs()[at] = e;
                    else pw.base[DANBO_PAIR_OVERFLOW_WORD] = 1;       // more visible pairs than the caller sized `work` for
                }
            }
        }
        __syncthreads();
    }
}

// Two pairs per lane (a warp = 64 pairs of the SAME bone): every aggregation-net weight is a warp-uniform load feeding
// 64 MACs.  With one pair per lane the kernel was bound by the load/store unit (one LSU pass per weight vector and per
// feature gather against four FMA pipes), not by FP32 throughput.
template <bool kOnePose>
__global__ void __launch_bounds__(128, 4)
pair_logits_kernel(const float* __restrict__ rays, int ray_stride, int S, const float* __restrict__ z,
                   const int* __restrict__ active_ids, const float* __restrict__ pose_skts,
                   const float* __restrict__ pose_vol, int rays_per_pose, int n_poses, FieldConsts fc, PairWork pw,
                   int pair_capacity, float* __restrict__ logits /* (n_rays*S, 24), visible entries only */) {
    constexpr int PP = 2;                                // pairs per lane
    const int lane = threadIdx.x & 31;
    // One pose (every render call): its 24 x 240 feature lines (23 KB) are staged in shared memory, where the 30
    // data-dependent taps per (pair, neighbour) cost ~1 pass each instead of one per touched 128-byte line of L1.
    __shared__ __align__(16) float vol_s[kOnePose ? DANBO_J * DANBO_VOL : 4];
    if (kOnePose) {
        const float4* src = reinterpret_cast<const float4*>(pose_vol);
        for (int i = threadIdx.x; i < DANBO_J * DANBO_VOL / 4; i += blockDim.x) reinterpret_cast<float4*>(vol_s)[i] = __ldg(src + i);
        __syncthreads();
    }
    const int* cnt = pw.count();
    // segments are padded to 32 pairs; a warp takes two consecutive 32-pair pieces of one bone's segment
    int n_chunks = 0;
    for (int j = 0; j < DANBO_J; ++j) n_chunks += (((cnt[j] + 31) >> 5) + PP - 1) / PP;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n_chunks; c += warps) {
        // bone of this chunk (warp-uniform) and the start of its segment in 32-pair units
        int j = 0, first = 0, seg32 = 0;
        for (;; ++j) {
            const int n32 = (cnt[j] + 31) >> 5, nc = (n32 + PP - 1) / PP;
            if (c < first + nc) break;
            first += nc; seg32 += n32;
        }
        const int n32_j = (cnt[j] + 31) >> 5;
        float px[PP], py[PP], pz[PP];
        int id[PP], pose[PP];
        bool live[PP];
#pragma unroll
        for (int u = 0; u < PP; ++u) {
            const int piece = (c - first) * PP + u;                  // 32-pair piece inside the bone's segment
            const int in_seg = piece * 32 + lane;
            const int at = (seg32 + piece) * 32 + lane;
            live[u] = piece < n32_j && in_seg < cnt[j] && at < pair_capacity;
            px[u] = py[u] = pz[u] = 0.f; id[u] = 0; pose[u] = 0;
            if (live[u]) {
                id[u] = active_ids[pw.pairs()[at]];
                const int n = id[u] / S;
                const float* r = rays + (size_t)n * ray_stride;
                const float zz = z[id[u]];
                px[u] = __fadd_rn(r[0], __fmul_rn(r[3], zz));
                py[u] = __fadd_rn(r[1], __fmul_rn(r[4], zz));
                pz[u] = __fadd_rn(r[2], __fmul_rn(r[5], zz));
                pose[u] = n / rays_per_pose; if (pose[u] >= n_poses) pose[u] = n_poses - 1;
            }
        }
        float mix[PP][DANBO_AGG_W];
#pragma unroll
        for (int u = 0; u < PP; ++u)
#pragma unroll
            for (int o = 0; o < DANBO_AGG_W; ++o) mix[u][o] = 0.f;
        uint32_t nb = kNbrMask[j];
        while (nb) {                                    // layer 0 of the bone's tree neighbours, mixed by adjacency
            const int k = __ffs(nb) - 1; nb &= nb - 1;
            const float adj = __ldg(fc.agg_adjw + j * DANBO_J + k) * __ldg(fc.agg_adj + j * DANBO_J + k);
            float h[PP][DANBO_FEAT];
#pragma unroll
            for (int u = 0; u < PP; ++u) {
                float x0, x1, x2;
                bone_coords(pose_skts + ((size_t)pose[u] * DANBO_J + k) * 16, fc.align + k * 16, fc.axis_scale + k * 3,
                            px[u], py[u], pz[u], x0, x1, x2);
                if (kOnePose) bone_features<true>(vol_s + k * DANBO_VOL, x0, x1, x2, h[u]);
                else bone_features<false>(pose_vol + ((size_t)pose[u] * DANBO_J + k) * DANBO_VOL, x0, x1, x2, h[u]);
            }
            const float4* w = reinterpret_cast<const float4*>(fc.agg_w0 + (size_t)k * DANBO_FEAT * DANBO_AGG_W);
#pragma unroll
            for (int i = 0; i < DANBO_FEAT; ++i) {
                float hi[PP];
#pragma unroll
                for (int u = 0; u < PP; ++u) hi[u] = h[u][i] * adj;
#pragma unroll
                for (int o4 = 0; o4 < DANBO_AGG_W / 4; ++o4) {
                    const float4 wv = __ldg(w + i * (DANBO_AGG_W / 4) + o4);
#pragma unroll
                    for (int u = 0; u < PP; ++u) {
                        mix[u][4 * o4 + 0] = fmaf(hi[u], wv.x, mix[u][4 * o4 + 0]); mix[u][4 * o4 + 1] = fmaf(hi[u], wv.y, mix[u][4 * o4 + 1]);
                        mix[u][4 * o4 + 2] = fmaf(hi[u], wv.z, mix[u][4 * o4 + 2]); mix[u][4 * o4 + 3] = fmaf(hi[u], wv.w, mix[u][4 * o4 + 3]);
                    }
                }
            }
        }
#pragma unroll
        for (int o = 0; o < DANBO_AGG_W; ++o) {
            const float b0 = __ldg(fc.agg_b0 + o);
#pragma unroll
            for (int u = 0; u < PP; ++u) mix[u][o] = fmaxf(mix[u][o] + b0, 0.f);
        }
        // layer 1 (32 -> 32) in two halves of 16 outputs to bound the live registers, then layer 2 (32 -> 1)
        float a[PP];
#pragma unroll
        for (int u = 0; u < PP; ++u) a[u] = __ldg(fc.agg_b2 + j);
        const float4* w1 = reinterpret_cast<const float4*>(fc.agg_w1 + (size_t)j * DANBO_AGG_W * DANBO_AGG_W);
#pragma unroll
        for (int oh = 0; oh < 2; ++oh) {
            float l1[PP][16];
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                const float b1 = __ldg(fc.agg_b1 + j * DANBO_AGG_W + oh * 16 + o);
#pragma unroll
                for (int u = 0; u < PP; ++u) l1[u][o] = b1;
            }
#pragma unroll
            for (int i = 0; i < DANBO_AGG_W; ++i) {
#pragma unroll
                for (int o4 = 0; o4 < 4; ++o4) {
                    const float4 wv = __ldg(w1 + i * (DANBO_AGG_W / 4) + oh * 4 + o4);
#pragma unroll
                    for (int u = 0; u < PP; ++u) {
                        l1[u][4 * o4 + 0] = fmaf(mix[u][i], wv.x, l1[u][4 * o4 + 0]); l1[u][4 * o4 + 1] = fmaf(mix[u][i], wv.y, l1[u][4 * o4 + 1]);
                        l1[u][4 * o4 + 2] = fmaf(mix[u][i], wv.z, l1[u][4 * o4 + 2]); l1[u][4 * o4 + 3] = fmaf(mix[u][i], wv.w, l1[u][4 * o4 + 3]);
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < 16; ++o) {
                const float w2 = __ldg(fc.agg_w2 + j * DANBO_AGG_W + oh * 16 + o);
#pragma unroll
                for (int u = 0; u < PP; ++u) a[u] = fmaf(fmaxf(l1[u][o], 0.f), w2, a[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < PP; ++u)
            if (live[u]) logits[(size_t)id[u] * DANBO_J + j] = a[u];
    }
}

__global__ void __launch_bounds__(128)
field_rows_kernel(const float* __restrict__ rays, int ray_stride, int n_rays, int S, const float* __restrict__ z,
                  const uint32_t* __restrict__ mask, const int* __restrict__ active_ids,
                  const int* __restrict__ active_count, int capacity, const float* __restrict__ pose_skts,
                  const float* __restrict__ pose_vol, int rays_per_pose, int n_poses, FieldConsts fc,
                  const float* __restrict__ logits, uint8_t* __restrict__ xtiles, int* __restrict__ row_ray,
                  float* __restrict__ hbar_out /* (rows,16) or null */,
                  __nv_bfloat16* __restrict__ x_rows /* (rows,208) row-major copy for the backward pass, or null */,
                  int agg_mode /* 0 sigmoid, 1 masked softmax */, const int* __restrict__ work) {
    int count = *active_count; if (count > capacity) count = capacity;
    const int total = n_rays * S;
    // pair list overflow (pair_bucket dropped pairs, so some logits were never written): every row of this call is
    // poisoned with NaN - the caller sees NaN pixels / a NaN loss instead of silently wrong densities
    const float poison = work[DANBO_PAIR_OVERFLOW_WORD] ? __int_as_float(0x7fc00000) : 0.f;
    // A lane owns a row, and a row's 16-byte pieces lie in different 128-byte lines of the tile image (one LSU pass per
    // lane per store, partial-line writes in L2).  Each warp therefore assembles its 32 rows of one 64-column chunk
    // (4 KB, contiguous in the image) in shared memory and sends it with a bulk copy; two buffers per warp.
    __shared__ __align__(128) uint8_t stage[4][2][4096];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int n_flush = 0;
    for (int base = blockIdx.x * blockDim.x; base < count; base += gridDim.x * blockDim.x) {
        if (base + warp * 32 >= count) break;           // this warp's rows (and those of its later tiles) do not exist
        const int e = base + threadIdx.x;
        const bool live = e < count;
        const int id = live ? active_ids[e] : total;
        float hbar[DANBO_FEAT];
#pragma unroll
        for (int i = 0; i < DANBO_FEAT; ++i) hbar[i] = 0.f;
        int n;
        if (id >= total) {
            n = id - total;                             // the ray's "no bone sees me" entry: h = 0
        } else {
            n = id / S;
            const float* r = rays + (size_t)n * ray_stride;
            const float zz = z[id];
            const float px = __fadd_rn(r[0], __fmul_rn(r[3], zz));
            const float py = __fadd_rn(r[1], __fmul_rn(r[4], zz));
            const float pz = __fadd_rn(r[2], __fmul_rn(r[5], zz));
            int pose = n / rays_per_pose; if (pose >= n_poses) pose = n_poses - 1;
            const float* skt = pose_skts + (size_t)pose * DANBO_J * 16;
            const float* vol = pose_vol + (size_t)pose * DANBO_J * DANBO_VOL;
            uint32_t m = mask[id];
            float amax = 0.f, inv_den = 0.f;
            if (agg_mode == 1) softmax_terms(logits + (size_t)id * DANBO_J, m, amax, inv_den);
            while (m) {
                const int j = __ffs(m) - 1; m &= m - 1;
                float x0, x1, x2, h[DANBO_FEAT];
                bone_coords(skt + j * 16, fc.align + j * 16, fc.axis_scale + j * 3, px, py, pz, x0, x1, x2);
                bone_features(vol + j * DANBO_VOL, x0, x1, x2, h);
                const float a = logits[(size_t)id * DANBO_J + j];
                const float p = agg_mode == 1 ? expf(a - amax) * inv_den                       // danbo.py:388-404
                                              : (1.f / (1.f + expf(-a))) * 1.002f - 0.001f;   // danbo.py:410, visible bone
#pragma unroll
                for (int i = 0; i < DANBO_FEAT; ++i) hbar[i] = fmaf(p, h[i], hbar[i]);
            }
        }
        if (poison != 0.f) {
#pragma unroll
            for (int i = 0; i < DANBO_FEAT; ++i) hbar[i] = poison;
        }
        if (live) row_ray[e] = n;
        if (hbar_out && live) {
            float4* ho = reinterpret_cast<float4*>(hbar_out + (size_t)e * 16);
            ho[0] = make_float4(hbar[0], hbar[1], hbar[2], hbar[3]);   ho[1] = make_float4(hbar[4], hbar[5], hbar[6], hbar[7]);
            ho[2] = make_float4(hbar[8], hbar[9], hbar[10], hbar[11]); ho[3] = make_float4(hbar[12], hbar[13], hbar[14], 0.f);
        }
        // ---- positional encoding (cutoff_embedder.py:62-73): [h, sin(2^0 h), cos(2^0 h), ..., cos(2^5 h)] -> bf16.
        // One sincos per feature; higher octaves by the double-angle recurrence (error ~1e-6, far below bf16).
        // Column of (octave f, fn, i) = 15 + 30 f + 15 fn + i; the row is emitted as 26 16-byte chunks.
        float sn[DANBO_FEAT], cs[DANBO_FEAT];
#pragma unroll
        for (int i = 0; i < DANBO_FEAT; ++i) sincosf(hbar[i], &sn[i], &cs[i]);
        const int tile = e >> 7, rr = e & 127;
        uint8_t* xt = xtiles + (size_t)tile * DANBO_X_TILE_BYTES;
        // 208 columns in order; emit 8 at a time.  Everything is unrolled so `col` is a compile-time constant.
        uint32_t pk[4];
#pragma unroll
        for (int col = 0; col < 208; ++col) {
            float v;
            if (col < DANBO_FEAT) v = hbar[col];
            else if (col < DANBO_X_COLS) {
                const int q = col - DANBO_FEAT, rem = q % 30;
                v = rem < DANBO_FEAT ? sn[rem] : cs[rem - DANBO_FEAT];
            } else v = 0.f;
            const uint32_t b = __bfloat16_as_ushort(__float2bfloat16_rn(v));
            if (col & 1) pk[(col & 7) >> 1] |= b << 16; else pk[(col & 7) >> 1] = b;
            if ((col & 7) == 7) {
                const int chunk = (col - 7) >> 6, unit = ((col - 7) & 63) >> 3;
                if (unit == 0) {                         // first piece of a chunk: its buffer must have been read out
                    if (n_flush >= 2) {
                        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                        __syncwarp();
                    }
                }
                const uint32_t dst = smem_u32(stage[warp][n_flush & 1]) + (uint32_t)(lane * 128 + ((unit ^ (lane & 7)) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                if (x_rows && live) *reinterpret_cast<uint4*>(x_rows + (size_t)e * 208 + (col - 7)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                if (unit == 7 || col == 207) {           // chunk complete (the last one holds 16 of its 64 columns)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        uint8_t* gdst = xt + chunk * 16384 + warp * 4096;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                                     ::"l"(gdst), "r"(smem_u32(stage[warp][n_flush & 1])), "r"(4096u) : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++n_flush;
                }
            }
            // after the last column of an octave (col == 14 + 30 (f+1)) advance every feature to the next octave
            if (col >= 44 && (col - 44) % 30 == 0 && col < 194) {
#pragma unroll
                for (int i = 0; i < DANBO_FEAT; ++i) { const float s2 = 2.f * sn[i] * cs[i], c2 = 1.f - 2.f * sn[i] * sn[i]; sn[i] = s2; cs[i] = c2; }
            }
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // shared memory is read out before the block ends
}

// ---------------------------------------------------------------------------------------------------------
// V1 folded into the view layer: b_ray[n] = W_v[:, 256:411] . [PE(rays_d) (27) ; frame code (128)] + b_v.
// nerf.py:252-279, embedding.py:86-108.  The frame-code part depends only on the camera index: it is computed once per
// code (code_table_kernel: n_codes + 1 rows, the last one for the mean code) and added to the per-ray 27-input part.
__global__ void __launch_bounds__(128)
code_table_kernel(const float* __restrict__ codes, const float* __restrict__ wv_ray, float* __restrict__ table) {
    const int ci = blockIdx.x, o = threadIdx.x;
    float acc = wv_ray[155 * 128 + o];
    for (int c = 0; c < 128; ++c) acc = fmaf(codes[(size_t)ci * 128 + c], wv_ray[(27 + c) * 128 + o], acc);
    table[ci * 128 + o] = acc;
}

// 32 rays per block of 128 threads; thread = output unit.
__global__ void __launch_bounds__(128)
ray_bias_kernel(const float* __restrict__ rays, int ray_stride, int n_rays, const int* __restrict__ cam_idx, int n_codes,
                const float* __restrict__ wv_ray /* (155,128) transposed view/code slice of W_v, then b_v (128) */,
                const float* __restrict__ table /* (n_codes+1,128) */, float* __restrict__ out /* (n_rays,128) */) {
    constexpr int RB = 32, VIN = 27;
    __shared__ __align__(16) float v[VIN][RB];          // inputs of 32 rays, ray-minor so one LDS.128 feeds 4 FMAs
    __shared__ int cam_s[RB];
    const int base = blockIdx.x * RB;
    for (int i = threadIdx.x; i < RB * VIN; i += blockDim.x) {
        const int rb = i % RB, c = i / RB;
        const int n = base + rb;
        float val = 0.f;
        if (n < n_rays) {
            const float* r = rays + (size_t)n * ray_stride + 3;
            if (c < 3) val = r[c];
            else { const int q = c - 3, f = q / 6, rem = q - 6 * f; const float x = r[rem % 3] * (float)(1 << f); val = rem < 3 ? sinf(x) : cosf(x); }
        }
        v[c][rb] = val;
    }
    if (threadIdx.x < RB) {
        const int n = base + threadIdx.x;
        int ci = (cam_idx && n < n_rays) ? cam_idx[n] : 0;
        cam_s[threadIdx.x] = ci < 0 ? n_codes : (ci > n_codes - 1 ? n_codes - 1 : ci);      // row n_codes holds the mean code
    }
    __syncthreads();
    const int o = threadIdx.x;
    float acc[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) acc[rb] = 0.f;
#pragma unroll 3
    for (int c = 0; c < VIN; ++c) {
        const float wc = __ldg(wv_ray + c * 128 + o);
        const float4* vr = reinterpret_cast<const float4*>(v[c]);
#pragma unroll
        for (int q = 0; q < RB / 4; ++q) {
            const float4 x = vr[q];
            acc[4 * q + 0] = fmaf(wc, x.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(wc, x.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(wc, x.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(wc, x.w, acc[4 * q + 3]);
        }
    }
#pragma unroll
    for (int rb = 0; rb < RB; ++rb)
        if (base + rb < n_rays) out[(size_t)(base + rb) * 128 + o] = acc[rb] + __ldg(table + cam_s[rb] * 128 + o);
}

}  // namespace danbo

using namespace danbo;

extern "C" int danbo_nearfar(const float* rays, int ray_stride, int n_rays, const float* pose_cyl, int cyl_stride,
                             const float* pose_skts, int rays_per_pose, int n_poses, const float* align,
                             const float* axis_scale, int seg_len, int use_box, float bound, float bound_hi, float* near_out,
                             float* far_out, double* seg_acc, int n_seg, unsigned char* p_valid,
                             unsigned char* v_valid, void* stream) {
    if (n_rays <= 0) return 0;
    if (rays_per_pose <= 0 || n_poses <= 0 || ray_stride < 8) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(seg_acc, 0, sizeof(double) * 4 * (size_t)n_seg, st);
    if (e != cudaSuccess) return (int)e;
    const int B = 128, G = (n_rays + B - 1) / B;
    nearfar_cyl_kernel<<<G, B, 0, st>>>(rays, ray_stride, n_rays, pose_cyl, cyl_stride, rays_per_pose, n_poses, seg_len,
                                        near_out, far_out, seg_acc);
    DANBO_CHECK_LAUNCH();
    FieldConsts fc{}; fc.align = align; fc.axis_scale = axis_scale;
    nearfar_finish_kernel<<<G, B, 0, st>>>(rays, ray_stride, n_rays, pose_skts, rays_per_pose, n_poses, fc, seg_len,
                                           seg_acc, use_box, bound, bound_hi, near_out, far_out, p_valid, v_valid);
    DANBO_CHECK_LAUNCH();
    return 0;
}

namespace danbo {
// field_mma.cu: the aggregation net as split-bf16 mma.sync products (opt-in through consts[10])
int launch_pair_logits_mma(const float* rays, int ray_stride, int S, const float* z, const int* active_ids,
                           const float* pose_skts, const float* pose_vol, int rays_per_pose, int n_poses,
                           const FieldConsts& fc, PairWork pw, int pair_capacity, const void* frags, float* logits,
                           int num_sms, cudaStream_t st);
}

static FieldConsts make_consts(const float* const* p) {
    FieldConsts fc;
    fc.align = p[0]; fc.axis_scale = p[1]; fc.agg_w0 = p[2]; fc.agg_adjw = p[3]; fc.agg_adj = p[4];
    fc.agg_b0 = p[5]; fc.agg_w1 = p[6]; fc.agg_b1 = p[7]; fc.agg_w2 = p[8]; fc.agg_b2 = p[9];
    return fc;
}

extern "C" int danbo_sample_mask(const float* rays, int ray_stride, int n_rays, int S, const float* near,
                                 const float* far, const float* t_vals, const float* t_rand, const float* z_in,
                                 float* z_out, const float* pose_skts, int rays_per_pose, int n_poses,
                                 const float* const* consts, unsigned int* mask_out, int* active_ids,
                                 int* active_count, int capacity, int append_empty, int lindisp, void* stream) {
    if (n_rays <= 0 || S <= 0) return 0;
    if (!z_in && (!near || !far || !t_vals || !z_out)) return -1;
    const long long total = (long long)n_rays * S;
    if (total + n_rays >= (1LL << 31)) return -2;
    const int B = kMaskBlock, G = (int)((total + B - 1) / B);
    if (S >= 16)
        sample_mask_kernel<true><<<G, B, 0, (cudaStream_t)stream>>>(rays, ray_stride, n_rays, S, near, far, t_vals, t_rand,
                                                                    z_in, z_out, pose_skts, rays_per_pose, n_poses,
                                                                    make_consts(consts), mask_out, active_ids,
                                                                    active_count, capacity, append_empty, lindisp);
    else
        sample_mask_kernel<false><<<G, B, 0, (cudaStream_t)stream>>>(rays, ray_stride, n_rays, S, near, far, t_vals, t_rand,
                                                                     z_in, z_out, pose_skts, rays_per_pose, n_poses,
                                                                     make_consts(consts), mask_out, active_ids,
                                                                     active_count, capacity, append_empty, lindisp);
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_field_agg(const float* rays, int ray_stride, int n_rays, int S, const float* z,
                               const unsigned int* mask, const int* active_ids, const int* active_count,
                               int capacity, const float* pose_skts, const float* pose_vol, int rays_per_pose,
                               int n_poses, const float* const* consts, void* xtiles, int* row_ray, float* logits,
                               float* hbar_out, void* x_rows, int* work, int pair_capacity, int num_sms, int agg_mode,
                               void* stream) {
    if (capacity <= 0) return 0;
    if (!logits || !work || pair_capacity < 32 || agg_mode < 0 || agg_mode > 1) return -1;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(work, 0, 64 * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    PairWork pw{work};
    const FieldConsts fc = make_consts(consts);
    const int total = n_rays * S;
    int blocks = (capacity + 255) / 256;
    if (blocks > num_sms * 8) blocks = num_sms * 8;
    pair_bucket_kernel<false><<<blocks, 256, 0, st>>>(mask, active_ids, active_count, capacity, total, pw, pair_capacity, agg_mode);
    DANBO_CHECK_LAUNCH();
    pair_bucket_kernel<true><<<blocks, 256, 0, st>>>(mask, active_ids, active_count, capacity, total, pw, pair_capacity, agg_mode);
    DANBO_CHECK_LAUNCH();
    int pblocks = (pair_capacity / 64 + 24 + 3) / 4;                  // two 32-pair pieces per warp (+ one odd piece per bone)
    if (pblocks > num_sms * 8) pblocks = num_sms * 8;
    const void* agg_frags = consts[10];                               // danbo_pack_agg_frags table, or NULL (FFMA kernel)
    if (agg_frags != nullptr) {
        const int rc = launch_pair_logits_mma(rays, ray_stride, S, z, active_ids, pose_skts, pose_vol, rays_per_pose,
                                              n_poses, fc, pw, pair_capacity, agg_frags, logits, num_sms, st);
        if (rc != 0) return rc;
    } else if (n_poses == 1)
        pair_logits_kernel<true><<<pblocks, 128, 0, st>>>(rays, ray_stride, S, z, active_ids, pose_skts, pose_vol, rays_per_pose,
                                                           n_poses, fc, pw, pair_capacity, logits);
    else
        pair_logits_kernel<false><<<pblocks, 128, 0, st>>>(rays, ray_stride, S, z, active_ids, pose_skts, pose_vol, rays_per_pose,
                                                            n_poses, fc, pw, pair_capacity, logits);
    DANBO_CHECK_LAUNCH();
    int rblocks = (capacity + 127) / 128;
    if (rblocks > num_sms * 16) rblocks = num_sms * 16;
    field_rows_kernel<<<rblocks, 128, 0, st>>>(rays, ray_stride, n_rays, S, z, mask, active_ids, active_count, capacity,
                                                pose_skts, pose_vol, rays_per_pose, n_poses, fc, logits,
                                                (uint8_t*)xtiles, row_ray, hbar_out, (__nv_bfloat16*)x_rows, agg_mode, work);
    DANBO_CHECK_LAUNCH();
    return 0;
}

extern "C" int danbo_ray_bias(const float* rays, int ray_stride, int n_rays, const int* cam_idx, const float* codes,
                              int n_codes, const float* wv_ray, float* table, float* out, void* stream) {
    if (n_rays <= 0) return 0;
    if (!table) return -1;
    code_table_kernel<<<n_codes + 1, 128, 0, (cudaStream_t)stream>>>(codes, wv_ray, table);
    DANBO_CHECK_LAUNCH();
    ray_bias_kernel<<<(n_rays + 31) / 32, 128, 0, (cudaStream_t)stream>>>(rays, ray_stride, n_rays, cam_idx, n_codes, wv_ray,
                                                                       table, out);
    DANBO_CHECK_LAUNCH();
    return 0;
}
