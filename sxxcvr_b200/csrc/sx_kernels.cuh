// sx_kernels.cuh -- sm_100a kernels for the SoapySX IQ sample path.
//
// Everything here is a streaming map: one output word depends on one input word (RX) or
// on one I/Q pair (TX).  There is no reuse and no contraction, so the bound is HBM
// bandwidth: 16 algorithmic bytes per frame for the CF32 paths (8 read + 8 written),
// 12 for the CS16 extensions, 24 for the fused loopback with the CF32 intermediate kept.
// The kernels therefore do exactly three things well: move full 128- or 256-bit vectors
// with every lane of a warp on consecutive addresses, keep enough of them in flight per
// SM to cover HBM latency, and stay out of L1/L2's way (no-allocate, evict-first).
//
// Three interchangeable schedules are provided for the equal-width conversions, selected
// by the host after measurement ("Kernel variants"):
//   vector128 / vector256 : persistent grid-stride kernel, UNROLL independent LDG.128 or
//                           LDG.256 per thread issued before the first use;
//   bulk                  : one elected thread drives cp.async.bulk (TMA) global->shared
//                           loads through an mbarrier ring, all threads convert
//                           shared->shared, one thread bulk-stores shared->global.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "sx_synth.h"

namespace sx {

// ---------------------------------------------------------------------------------------
// Register packs and streaming global access
// ---------------------------------------------------------------------------------------
template <int WORDS> struct Pack {
    uint32_t w[WORDS];
};

// Loads do not use .nc: the CF32 conversions may run in place (src == dest), which the
// non-coherent path does not allow.  L1::no_allocate keeps once-touched lines out of L1;
// the 256-bit forms also carry L2::evict_first (only they can, without a policy operand).
template <int BYTES> __device__ __forceinline__ Pack<BYTES / 4> ld_stream(const void *p);
template <int BYTES> __device__ __forceinline__ void st_stream(void *p, const Pack<BYTES / 4> &v);

template <> __device__ __forceinline__ Pack<1> ld_stream<4>(const void *p)
{
    Pack<1> r;
    asm volatile("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(r.w[0]) : "l"(p) : "memory");
    return r;
}
template <> __device__ __forceinline__ Pack<2> ld_stream<8>(const void *p)
{
    Pack<2> r;
    asm volatile("ld.global.L1::no_allocate.v2.b32 {%0,%1}, [%2];"
                 : "=r"(r.w[0]), "=r"(r.w[1])
                 : "l"(p)
                 : "memory");
    return r;
}
template <> __device__ __forceinline__ Pack<4> ld_stream<16>(const void *p)
{
    Pack<4> r;
    asm volatile("ld.global.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3])
                 : "l"(p)
                 : "memory");
    return r;
}
template <> __device__ __forceinline__ Pack<8> ld_stream<32>(const void *p)
{
    Pack<8> r;
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3]), "=r"(r.w[4]),
                   "=r"(r.w[5]), "=r"(r.w[6]), "=r"(r.w[7])
                 : "l"(p)
                 : "memory");
    return r;
}

template <> __device__ __forceinline__ void st_stream<4>(void *p, const Pack<1> &v)
{
    asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(p), "r"(v.w[0]) : "memory");
}
template <> __device__ __forceinline__ void st_stream<8>(void *p, const Pack<2> &v)
{
    asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" ::"l"(p), "r"(v.w[0]),
                 "r"(v.w[1])
                 : "memory");
}
template <> __device__ __forceinline__ void st_stream<16>(void *p, const Pack<4> &v)
{
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.w[0]),
                 "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3])
                 : "memory");
}
template <> __device__ __forceinline__ void st_stream<32>(void *p, const Pack<8> &v)
{
    asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p),
                 "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]), "r"(v.w[4]), "r"(v.w[5]),
                 "r"(v.w[6]), "r"(v.w[7])
                 : "memory");
}

// ---------------------------------------------------------------------------------------
// The conversions.  Each Op converts FR whole frames held in registers.
// kSrcWords / kDstWords = 32-bit words per frame on each side.
// ---------------------------------------------------------------------------------------

// RX S32 -> CF32.  Reference SoapySX.cpp:107-110: dest = 2^-31 * (float)src.
// cvt.rn.f32.s32 rounds to nearest-even exactly like the CPU's int->float, and the
// multiply by a power of two is exact (no subnormals: the smallest nonzero result is
// 2^-31), so contraction or reassociation cannot change a bit.
struct RxCf32 {
    static constexpr int kSrcWords = 2, kDstWords = 2;
    template <int FR>
    __device__ __forceinline__ static void apply(const Pack<2 * FR> &in, Pack<2 * FR> &out, float)
    {
#pragma unroll
        for (int i = 0; i < 2 * FR; i++)
            out.w[i] = __float_as_uint(__int2float_rn(int(in.w[i])) * 4.656612873077392578125e-10f);
    }
};

// One TX component.  Reference SoapySX.cpp:124-125, :130-131:
//     v = (int32)(2^31 * max(min(f, 1), -1)) & ~3
// cvt.rzi.s32.f32 truncates toward zero AND saturates, with NaN -> 0.  Saturation makes the
// clamp redundant: f >= 1 gives 2^31*f >= 2^31 -> INT32_MAX, exactly what the clamped value
// 2^31 gives; f <= -1 gives <= -2^31 -> INT32_MIN, exactly the clamped -2^31; the product
// itself is exact (power of two) and +-inf saturate the same way.  NaN passes through the
// reference's std::min/std::max (they return their first argument when the comparison is
// false) and converts to 0 on the reference's ARM target, as here.
__device__ __forceinline__ uint32_t tx_component(float f)
{
    return uint32_t(__float2int_rz(f * 2147483648.0f)) & 0xFFFFFFFCu;
}

// TX-enable decision, reference SoapySX.cpp:132: fi*fi + fq*fq >= thr2 on the un-clamped
// inputs, three separately rounded operations.  The _rn intrinsics are never contracted
// into an FMA, which matters: within 2 ulp of the threshold circle a fused evaluation
// disagrees on 2.3 % of points.  NaN compares false.
__device__ __forceinline__ uint32_t tx_enable_bits(float fi, float fq, float thr2)
{
    float mag2 = __fadd_rn(__fmul_rn(fi, fi), __fmul_rn(fq, fq));
    return (mag2 >= thr2) ? 3u : 0u;
}

// TX CF32 -> S32.
struct TxCf32 {
    static constexpr int kSrcWords = 2, kDstWords = 2;
    template <int FR>
    __device__ __forceinline__ static void apply(const Pack<2 * FR> &in, Pack<2 * FR> &out,
                                                 float thr2)
    {
#pragma unroll
        for (int n = 0; n < FR; n++) {
            float fi = __uint_as_float(in.w[2 * n]), fq = __uint_as_float(in.w[2 * n + 1]);
            out.w[2 * n] = tx_component(fi) | tx_enable_bits(fi, fq, thr2);
            out.w[2 * n + 1] = tx_component(fq);
        }
    }
};

// EXTENSION (no reference): RX S32 -> CS16, out = (int16)(word >> 16).  One PRMT per frame:
// bytes 2,3 of I and bytes 2,3 of Q.
struct RxCs16 {
    static constexpr int kSrcWords = 2, kDstWords = 1;
    template <int FR>
    __device__ __forceinline__ static void apply(const Pack<2 * FR> &in, Pack<FR> &out, float)
    {
#pragma unroll
        for (int n = 0; n < FR; n++)
            out.w[n] = __byte_perm(in.w[2 * n], in.w[2 * n + 1], 0x7632);
    }
};

// EXTENSION (no reference): TX CS16 -> S32, word = s << 16; flag on f = s * 2^-15.
struct TxCs16 {
    static constexpr int kSrcWords = 1, kDstWords = 2;
    template <int FR>
    __device__ __forceinline__ static void apply(const Pack<FR> &in, Pack<2 * FR> &out, float thr2)
    {
#pragma unroll
        for (int n = 0; n < FR; n++) {
            uint32_t w = in.w[n];
            float fi = __int2float_rn(int(short(w & 0xFFFFu))) * 3.0517578125e-05f;
            float fq = __int2float_rn(int(w) >> 16) * 3.0517578125e-05f;
            out.w[2 * n] = (w << 16) | tx_enable_bits(fi, fq, thr2);
            out.w[2 * n + 1] = w & 0xFFFF0000u;
        }
    }
};

// EXTENSION (no reference): 16-bit I2S slots, frame = [I:int16][Q:int16] in one word.
// RX: f = s * 2^-15 (exact).
struct RxS16Cf32 {
    static constexpr int kSrcWords = 1, kDstWords = 2;
    template <int FR>
    __device__ __forceinline__ static void apply(const Pack<FR> &in, Pack<2 * FR> &out, float)
    {
#pragma unroll
        for (int n = 0; n < FR; n++) {
            uint32_t w = in.w[n];
            out.w[2 * n] = __float_as_uint(__int2float_rn(int(short(w & 0xFFFFu))) * 3.0517578125e-05f);
            out.w[2 * n + 1] = __float_as_uint(__int2float_rn(int(w) >> 16) * 3.0517578125e-05f);
        }
    }
};

// TX: v = trunc(2^15 * f) saturated to int16, low two bits cleared; flag bits as in TxCf32.
struct TxCf32S16 {
    static constexpr int kSrcWords = 2, kDstWords = 1;
    __device__ __forceinline__ static uint32_t component(float f)
    {
        int v = __float2int_rz(f * 32768.0f); // saturates at the int32 rails, NaN -> 0
        v = v > 32767 ? 32767 : (v < -32768 ? -32768 : v);
        return uint32_t(v) & 0xFFFCu;
    }
    template <int FR>
    __device__ __forceinline__ static void apply(const Pack<2 * FR> &in, Pack<FR> &out, float thr2)
    {
#pragma unroll
        for (int n = 0; n < FR; n++) {
            float fi = __uint_as_float(in.w[2 * n]), fq = __uint_as_float(in.w[2 * n + 1]);
            out.w[n] = (component(fi) | tx_enable_bits(fi, fq, thr2)) | (component(fq) << 16);
        }
    }
};

// ---------------------------------------------------------------------------------------
// vector128 / vector256: persistent grid-stride streaming kernel
// ---------------------------------------------------------------------------------------
struct StreamArgs {
    const char *src; // first frame
    char *dst;       // first frame
    uint64_t total;  // frames in the call
    uint64_t head;   // frames before the vector-aligned middle (done frame-by-frame)
    uint64_t nvec;   // vector accesses in the middle, FR frames each
    float thr2;
};

// FR frames per access (2 -> 128-bit on the 8-byte/frame side, 4 -> 256-bit).
// UNROLL accesses per thread are loaded before any is used: UNROLL*FR*8 bytes in flight
// per thread is what covers HBM latency, not occupancy alone.  BLOCK is a compile-time
// constant so that the UNROLL addresses are one base register plus immediates and ptxas
// can issue every load of a tile back to back.
template <class Op, int FR, int UNROLL, int BLOCK>
__global__ void __launch_bounds__(BLOCK) stream_convert_kernel(const StreamArgs a)
{
    constexpr int SB = Op::kSrcWords * 4 * FR, DB = Op::kDstWords * 4 * FR;
    constexpr uint64_t TILE = uint64_t(BLOCK) * UNROLL;
    const char *src = a.src + a.head * (Op::kSrcWords * 4);
    char *dst = a.dst + a.head * (Op::kDstWords * 4);

    for (uint64_t base = uint64_t(blockIdx.x) * TILE; base < a.nvec;
         base += uint64_t(gridDim.x) * TILE) {
        Pack<Op::kSrcWords * FR> in[UNROLL];
        Pack<Op::kDstWords * FR> out[UNROLL];
        const uint64_t first = base + threadIdx.x;
        const char *sp = src + first * SB;
        char *dp = dst + first * DB;
        if (base + TILE <= a.nvec) { // full tile: no per-access predicate
#pragma unroll
            for (int j = 0; j < UNROLL; j++)
                in[j] = ld_stream<SB>(sp + size_t(j) * BLOCK * SB);
#pragma unroll
            for (int j = 0; j < UNROLL; j++)
                Op::template apply<FR>(in[j], out[j], a.thr2);
#pragma unroll
            for (int j = 0; j < UNROLL; j++)
                st_stream<DB>(dp + size_t(j) * BLOCK * DB, out[j]);
        } else {
#pragma unroll
            for (int j = 0; j < UNROLL; j++) {
                if (first + uint64_t(j) * BLOCK < a.nvec) {
                    in[j] = ld_stream<SB>(sp + size_t(j) * BLOCK * SB);
                    Op::template apply<FR>(in[j], out[j], a.thr2);
                    st_stream<DB>(dp + size_t(j) * BLOCK * DB, out[j]);
                }
            }
        }
    }

    // Edge frames around the aligned middle: fewer than 2*FR of them, one frame per thread.
    if (blockIdx.x == 0) {
        const uint64_t mid = a.nvec * FR;
        const uint64_t nedge = a.total - mid;
        if (threadIdx.x < nedge) {
            uint64_t f = threadIdx.x < a.head ? threadIdx.x : mid + threadIdx.x;
            Pack<Op::kSrcWords> in1 = ld_stream<Op::kSrcWords * 4>(a.src + f * (Op::kSrcWords * 4));
            Pack<Op::kDstWords> out1;
            Op::template apply<1>(in1, out1, a.thr2);
            st_stream<Op::kDstWords * 4>(a.dst + f * (Op::kDstWords * 4), out1);
        }
    }
}

// Fallback for buffers that are only 4-byte aligned: word accesses, one frame per thread
// per step.  Correctness path, not a performance path.
template <class Op>
__global__ void word_convert_kernel(const char *src, char *dst, uint64_t total, float thr2)
{
    for (uint64_t f = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; f < total;
         f += uint64_t(gridDim.x) * blockDim.x) {
        Pack<Op::kSrcWords> in;
        Pack<Op::kDstWords> out;
#pragma unroll
        for (int i = 0; i < Op::kSrcWords; i++)
            in.w[i] = ld_stream<4>(src + (f * Op::kSrcWords + i) * 4).w[0];
        Op::template apply<1>(in, out, thr2);
#pragma unroll
        for (int i = 0; i < Op::kDstWords; i++) {
            Pack<1> o;
            o.w[0] = out.w[i];
            st_stream<4>(dst + (f * Op::kDstWords + i) * 4, o);
        }
    }
}

// ---------------------------------------------------------------------------------------
// bulk: cp.async.bulk (TMA) staged through shared memory
// ---------------------------------------------------------------------------------------
namespace bulk {

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return uint32_t(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "WAIT_%=:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra DONE_%=;\n"
                 "bra WAIT_%=;\n"
                 "DONE_%=:\n"
                 "}" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
// L2 eviction policy for a once-touched stream.  Which one is best is an empirical question
// (it changes how the L2 batches write-backs to HBM), so the kernel takes it as an argument:
// 0 evict_first, 1 evict_normal, 2 evict_last, 3 evict_unchanged.
__device__ __forceinline__ uint64_t make_policy(int kind)
{
    uint64_t pol;
    switch (kind) {
    case 1: asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol)); break;
    case 2: asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol)); break;
    case 3: asm volatile("createpolicy.fractional.L2::evict_unchanged.b64 %0, 1.0;" : "=l"(pol)); break;
    default: asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); break;
    }
    return pol;
}
__device__ __forceinline__ void load_g2s(void *smem_dst, const void *gsrc, uint32_t bytes,
                                         uint64_t *bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
                 "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void store_s2g(void *gdst, const void *smem_src, uint32_t bytes,
                                          uint64_t pol)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void commit_group()
{
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N> __device__ __forceinline__ void wait_group_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void wait_group_all()
{
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

} // namespace bulk

struct BulkArgs {
    const char *src;  // 16-byte aligned, start of the aligned middle (SKEW: 8 bytes before it)
    char *dst;        // 16-byte aligned
    uint64_t nframes; // frames in the middle; nframes * frame bytes is a multiple of 16 on both sides
    float thr2;
    int load_policy;  // see bulk::make_policy
    int store_policy;
    int contiguous;   // 0: tile t belongs to CTA t mod grid (round robin); 1: each CTA owns one contiguous range
};

// Persistent CTAs; tile t of TILE frames belongs to CTA (t mod gridDim.x).  Per CTA a ring
// of STAGES input buffers is kept full by thread 0 (bulk loads complete on an mbarrier);
// all threads convert stage s from the input buffer into the matching output buffer with
// 128-bit shared accesses; thread 0 then bulk-stores the output buffer and refills the
// input buffer with the tile STAGES ahead.  An output buffer is rewritten only after the
// bulk store that last read it has drained (wait_group.read).
// Dynamic shared memory: STAGES * (TILE * (src + dst frame bytes) + skew pad) + STAGES * 8.
//
// SKEW (8-byte frames on both sides only): the source is one frame out of step with the
// destination -- 8 modulo 16 where the destination is 16-byte aligned -- which bulk copies
// cannot express.  Then a.src is the source address rounded DOWN to 16 bytes, every tile
// loads 16 bytes more than it needs (8 before, 8 after), and the conversion reads shared
// memory 8 bytes in.  The host sizes head and tail so that this superset never leaves the
// caller's range (launch_bulk).
template <class Op, int TILE, int STAGES, bool SKEW = false>
__global__ void bulk_convert_kernel(const BulkArgs a)
{
    constexpr int SFB = Op::kSrcWords * 4, DFB = Op::kDstWords * 4; // frame bytes
    static_assert(!SKEW || (SFB == 8 && DFB == 8), "the skewed variant is for 8-byte frames on both sides");
    constexpr int PAD = SKEW ? 16 : 0;                 // extra bytes loaded per tile
    constexpr size_t IN_STAGE = size_t(TILE) * SFB + PAD;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *in_buf = smem;
    unsigned char *out_buf = smem + size_t(STAGES) * IN_STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(out_buf + size_t(STAGES) * TILE * DFB);

    const uint64_t ntiles = (a.nframes + TILE - 1) / TILE;
    // Tile i of this CTA is global tile first + i * stride.
    uint64_t first, stride, mine;
    if (a.contiguous) {
        const uint64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
        first = uint64_t(blockIdx.x) * per;
        stride = 1;
        mine = first >= ntiles ? 0 : (ntiles - first < per ? ntiles - first : per);
    } else {
        first = blockIdx.x;
        stride = gridDim.x;
        mine = first >= ntiles ? 0 : (ntiles - first + stride - 1) / stride;
    }
    if (mine == 0)
        return;
    const uint64_t pol = bulk::make_policy(a.load_policy), pol_store = bulk::make_policy(a.store_policy);

    auto tile_frames = [&](uint64_t i) -> uint32_t {
        uint64_t t = first + i * stride;
        uint64_t left = a.nframes - t * TILE;
        return uint32_t(left < uint64_t(TILE) ? left : uint64_t(TILE));
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++)
            bulk::mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x == 0) {
        for (uint64_t i = 0; i < uint64_t(STAGES) && i < mine; i++) {
            uint32_t nf = tile_frames(i);
            bulk::mbar_expect_tx(&full[i], nf * SFB + PAD);
            bulk::load_g2s(in_buf + i * IN_STAGE, a.src + (first + i * stride) * TILE * SFB,
                           nf * SFB + PAD, &full[i], pol);
        }
    }

    for (uint64_t i = 0; i < mine; i++) {
        const int s = int(i % STAGES);
        const uint32_t parity = uint32_t(i / STAGES) & 1u;
        const uint32_t nf = tile_frames(i);

        // out_buf[s] was last read by the bulk store of tile i-STAGES.
        if (threadIdx.x == 0)
            bulk::wait_group_read<STAGES - 1>();
        __syncthreads();

        bulk::mbar_wait(&full[s], parity);

        // 2 frames per access on the 8-byte/frame side; nf is even for every tile because
        // TILE is even and the middle is a whole number of 16-byte units.
        const unsigned char *ib = in_buf + size_t(s) * IN_STAGE + (SKEW ? 8 : 0);
        unsigned char *ob = out_buf + size_t(s) * TILE * DFB;
        constexpr int FR = (Op::kSrcWords == 1 || Op::kDstWords == 1) ? 4 : 2;
        for (uint32_t v = threadIdx.x; v * FR < nf; v += blockDim.x) {
            Pack<Op::kSrcWords * FR> in;
            Pack<Op::kDstWords * FR> out;
            if constexpr (SKEW) { // 8 bytes off a 16-byte boundary: frame-wide shared loads
                const uint2 *ip = reinterpret_cast<const uint2 *>(ib + size_t(v) * FR * SFB);
#pragma unroll
                for (int q = 0; q < FR; q++) {
                    uint2 t = ip[q];
                    in.w[2 * q] = t.x, in.w[2 * q + 1] = t.y;
                }
            } else {
                const uint4 *ip = reinterpret_cast<const uint4 *>(ib + size_t(v) * FR * SFB);
#pragma unroll
                for (int q = 0; q < Op::kSrcWords * FR / 4; q++) {
                    uint4 t = ip[q];
                    in.w[4 * q] = t.x, in.w[4 * q + 1] = t.y, in.w[4 * q + 2] = t.z, in.w[4 * q + 3] = t.w;
                }
            }
            Op::template apply<FR>(in, out, a.thr2);
            uint4 *op = reinterpret_cast<uint4 *>(ob + size_t(v) * FR * DFB);
#pragma unroll
            for (int q = 0; q < Op::kDstWords * FR / 4; q++)
                op[q] = make_uint4(out.w[4 * q], out.w[4 * q + 1], out.w[4 * q + 2], out.w[4 * q + 3]);
        }
        bulk::fence_async_smem(); // make the generic-proxy writes visible to the bulk store
        __syncthreads();

        if (threadIdx.x == 0) {
            bulk::store_s2g(a.dst + (first + i * stride) * TILE * DFB, ob, nf * DFB, pol_store);
            bulk::commit_group();
            uint64_t nxt = i + STAGES;
            if (nxt < mine) {
                uint32_t nnf = tile_frames(nxt);
                bulk::mbar_expect_tx(&full[s], nnf * SFB + PAD);
                bulk::load_g2s(in_buf + size_t(s) * IN_STAGE,
                               a.src + (first + nxt * stride) * TILE * SFB, nnf * SFB + PAD, &full[s], pol);
            }
        }
    }
    if (threadIdx.x == 0)
        bulk::wait_group_all();
}

// ---------------------------------------------------------------------------------------
// Fused repeater: RX-convert, keep the CF32 intermediate (optional), TX-convert.
// ---------------------------------------------------------------------------------------
struct LoopbackArgs {
    const char *i2s_in;
    char *cf32; // may be null
    char *i2s_out;
    uint64_t nvec; // 2-frame (16-byte) accesses; all three buffers 16-byte aligned
    uint64_t total; // frames; an odd last frame is done by block 0
    float thr2;
};

template <int UNROLL>
__global__ void loopback_kernel(const LoopbackArgs a)
{
    const uint64_t tile = uint64_t(blockDim.x) * UNROLL;
    for (uint64_t base = uint64_t(blockIdx.x) * tile; base < a.nvec;
         base += uint64_t(gridDim.x) * tile) {
        Pack<4> in[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            uint64_t v = base + threadIdx.x + uint64_t(j) * blockDim.x;
            if (v < a.nvec)
                in[j] = ld_stream<16>(a.i2s_in + v * 16);
        }
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            uint64_t v = base + threadIdx.x + uint64_t(j) * blockDim.x;
            if (v < a.nvec) {
                Pack<4> mid, out;
                RxCf32::apply<2>(in[j], mid, 0.0f);
                TxCf32::apply<2>(mid, out, a.thr2);
                if (a.cf32)
                    st_stream<16>(a.cf32 + v * 16, mid);
                st_stream<16>(a.i2s_out + v * 16, out);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && (a.total & 1)) {
        uint64_t f = a.total - 1;
        Pack<2> in1 = ld_stream<8>(a.i2s_in + f * 8), mid, out;
        RxCf32::apply<1>(in1, mid, 0.0f);
        TxCf32::apply<1>(mid, out, a.thr2);
        if (a.cf32)
            st_stream<8>(a.cf32 + f * 8, mid);
        st_stream<8>(a.i2s_out + f * 8, out);
    }
}

// ---------------------------------------------------------------------------------------
// Batched blocks: one warp per block (small blocks), or a slice of CTAs per block.
// Mirrors sxgpu_block in include/sxgpu.h.
// ---------------------------------------------------------------------------------------
struct BlockDesc {
    const char *src;
    char *dst;
    uint64_t length;
    float thr2;
    uint32_t reserved;
};

// Converts frames [lo, hi) of one block with `nthreads` cooperating threads, of which this
// is number `tid`.  Uses 128-bit accesses when both sides of the block allow it.
template <class Op>
__device__ __forceinline__ void convert_span(const BlockDesc &b, uint64_t lo, uint64_t hi,
                                             uint32_t tid, uint32_t nthreads)
{
    constexpr int SFB = Op::kSrcWords * 4, DFB = Op::kDstWords * 4;
    constexpr int FR = 16 / (SFB < DFB ? SFB : DFB); // frames per access: narrow side = 128 bit
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(b.src) + lo * SFB) % (SFB * FR) == 0) &&
                        ((reinterpret_cast<uintptr_t>(b.dst) + lo * DFB) % (DFB * FR) == 0);
    uint64_t done = lo;
    if (vec_ok) {
        uint64_t nvec = (hi - lo) / FR;
        // Four loads in flight per thread before the first store: a 256-frame period is four
        // vectors per lane of a warp, i.e. one round trip to L2/HBM instead of four in a row.
        constexpr int U = 4;
        uint64_t v = tid;
        for (; v + uint64_t(U - 1) * nthreads < nvec; v += uint64_t(U) * nthreads) {
            Pack<Op::kSrcWords * FR> in[U];
#pragma unroll
            for (int u = 0; u < U; u++)
                in[u] = ld_stream<SFB * FR>(b.src + (lo + (v + uint64_t(u) * nthreads) * FR) * SFB);
#pragma unroll
            for (int u = 0; u < U; u++) {
                Pack<Op::kDstWords * FR> out;
                Op::template apply<FR>(in[u], out, b.thr2);
                st_stream<DFB * FR>(b.dst + (lo + (v + uint64_t(u) * nthreads) * FR) * DFB, out);
            }
        }
        for (; v < nvec; v += nthreads) {
            Pack<Op::kSrcWords * FR> in = ld_stream<SFB * FR>(b.src + (lo + v * FR) * SFB);
            Pack<Op::kDstWords * FR> out;
            Op::template apply<FR>(in, out, b.thr2);
            st_stream<DFB * FR>(b.dst + (lo + v * FR) * DFB, out);
        }
        done = lo + nvec * FR;
    }
    for (uint64_t f = done + tid; f < hi; f += nthreads) {
        Pack<Op::kSrcWords> in = ld_stream<SFB>(b.src + f * SFB);
        Pack<Op::kDstWords> out;
        Op::template apply<1>(in, out, b.thr2);
        st_stream<DFB>(b.dst + f * DFB, out);
    }
}

// Small blocks: warp w of the grid takes blocks w, w + nwarps, ...
template <class Op>
__global__ void batch_warp_kernel(const BlockDesc *__restrict__ blocks, uint32_t nblocks)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t warps_per_cta = blockDim.x >> 5;
    const uint64_t nwarps = uint64_t(gridDim.x) * warps_per_cta;
    for (uint64_t b = uint64_t(blockIdx.x) * warps_per_cta + (threadIdx.x >> 5); b < nblocks;
         b += nwarps) {
        const BlockDesc d = blocks[b];
        convert_span<Op>(d, 0, d.length, lane, 32);
    }
}

// Large blocks: gridDim.y slices per block, each slice a contiguous range rounded to
// 32-frame units so that vector alignment carries over from the block start.
template <class Op>
__global__ void batch_slice_kernel(const BlockDesc *__restrict__ blocks, uint32_t nblocks)
{
    for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const BlockDesc d = blocks[b];
        uint64_t units = (d.length + 31) / 32;
        uint64_t per = (units + gridDim.y - 1) / gridDim.y;
        uint64_t lo = uint64_t(blockIdx.y) * per * 32;
        uint64_t hi = lo + per * 32;
        if (hi > d.length)
            hi = d.length;
        if (lo < hi)
            convert_span<Op>(d, lo, hi, threadIdx.x, blockDim.x);
    }
}

// Mid-size and large blocks, one chunk per CTA: CTA i takes chunk (i mod chunks_per_block) of
// block (i / chunks_per_block), so that the hardware hands the CTAs out block after block, chunk
// after chunk -- the order the blocks usually lie in memory.  chunks_per_block is sized for the
// longest block; CTAs past a shorter block's end leave at once, and the last CTA of a block takes
// all that is left of it.  A chunk is four accesses per thread (all four loads in flight before
// the first store, see convert_span).
template <class Op> __host__ __device__ constexpr uint64_t batch_direct_chunk()
{
    constexpr int narrow = Op::kSrcWords < Op::kDstWords ? Op::kSrcWords : Op::kDstWords;
    return uint64_t(256) * 4 * (16 / (narrow * 4));
}
template <class Op>
__global__ void __launch_bounds__(256) batch_direct_kernel(const BlockDesc *__restrict__ blocks, uint32_t nblocks,
                                                           uint32_t chunks_per_block)
{
    const uint32_t b = blockIdx.x / chunks_per_block;
    if (b >= nblocks)
        return;
    const BlockDesc d = blocks[b];
    const uint32_t c = blockIdx.x - b * chunks_per_block;
    const uint64_t lo = uint64_t(c) * batch_direct_chunk<Op>();
    if (lo >= d.length)
        return;
    // The last CTA of a block takes whatever is left, so that a block longer than the caller said
    // (max_length) is still converted whole, only slowly.
    const uint64_t hi = (c + 1 < chunks_per_block && lo + batch_direct_chunk<Op>() < d.length) ? lo + batch_direct_chunk<Op>()
                                                                                              : d.length;
    convert_span<Op>(d, lo, hi, threadIdx.x, 256);
}

// ---------------------------------------------------------------------------------------
// Batched blocks on the bulk-async schedule: many mid-size blocks (BASELINE config 5's small
// end: 1 MiB - 64 MiB each) in ONE launch at the large-block kernel's throughput.
//
// Every block is cut into tiles of TILE frames; tile_start[b] = number of tiles in blocks
// 0..b-1 (exclusive prefix sum, tile_start[nblocks] = total).  Persistent CTAs walk the tiles
// of all blocks the way bulk_convert_kernel walks the tiles of one block,
// and run the same pipeline (bulk loads into a STAGES-deep shared-memory ring on mbarriers,
// shared -> shared conversion by all threads, bulk stores gated by wait_group.read), except
// that a CTA owns one contiguous range of tiles: the producer thread finds its first tile's block
// by binary search, keeps that block's descriptor in registers, has the next block's already
// requested, and so reads the descriptor list only when a block ends -- a per-tile lookup (two
// dependent loads, ~1 us from L2) on the thread the whole CTA waits for cost 10-80 % of the
// throughput (profiles/r02_summary.md).  It leaves a small record per stage for the consumers.
//
// A block whose buffers are not 16-byte aligned cannot be moved by bulk copies: its tiles are
// converted with ordinary vector accesses by the same CTA ("direct" tiles), in the same pass.
// An odd last frame (8 bytes past the 16-byte units) is converted directly by one thread.
// ---------------------------------------------------------------------------------------
struct BatchBulkArgs {
    const BlockDesc *blocks;
    const unsigned long long *tile_start; // [nblocks + 1] in device memory, or null: summed in the kernel
    uint32_t nblocks;
    int load_policy, store_policy;
};

// Lists of up to this many blocks need no tile_start from outside: every CTA sums the lengths
// itself into shared memory (16 KiB), which saves a kernel launch and a scratch allocation per
// batch -- together more than the conversion of a small batch takes.
constexpr uint32_t kBatchLocalBlocks = 2048;

struct TileRecord {
    const char *src; // first frame of the tile
    char *dst;
    uint32_t bulk_frames; // frames moved by the bulk copies (even); 0 = direct tile
    uint32_t frames;      // frames in the tile, bulk + direct remainder
    float thr2;
    uint32_t pad;
};

// Producer-side cursor over the blocks of a batch: the block the current tile belongs to, held
// in registers, and the NEXT non-trivial facts (next block's descriptor and end) already
// requested, so that crossing a block boundary costs no round trip to memory at that moment.
struct BatchCursor {
    uint32_t blk;                 // current block
    unsigned long long first, end; // its tiles are [first, end)
    BlockDesc desc;
    BlockDesc next_desc;          // blocks[blk + 1] (prefetched; junk past the last block)
    unsigned long long next_end;  // tile_start[blk + 2]
};

__device__ __forceinline__ void batch_cursor_prefetch(const BatchBulkArgs &a, const unsigned long long *ts, BatchCursor &c)
{
    if (c.blk + 1 < a.nblocks) {
        c.next_desc = a.blocks[c.blk + 1];
        c.next_end = ts[c.blk + 2];
    }
}

__device__ __forceinline__ void batch_cursor_seek(const BatchBulkArgs &a, const unsigned long long *ts, BatchCursor &c,
                                                  unsigned long long t)
{
    // binary search: the last block whose first tile is <= t (steps over empty blocks)
    uint32_t lo = 0, hi = a.nblocks - 1;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo + 1) / 2;
        if (ts[mid] <= t)
            lo = mid;
        else
            hi = mid - 1;
    }
    c.blk = lo;
    c.first = ts[lo];
    c.end = ts[lo + 1];
    c.desc = a.blocks[lo];
    batch_cursor_prefetch(a, ts, c);
}

template <int TILE>
__device__ __forceinline__ void batch_tile_record(const BatchBulkArgs &a, const unsigned long long *ts, unsigned long long t,
                                                  BatchCursor &c, TileRecord &rec, int src_frame_bytes, int dst_frame_bytes)
{
    while (t >= c.end) { // next block (tiles arrive in increasing order); empty blocks fall through
        c.blk++;
        c.first = c.end;
        c.end = c.next_end;
        c.desc = c.next_desc;
        batch_cursor_prefetch(a, ts, c);
    }
    const uint64_t lo = (t - c.first) * uint64_t(TILE);
    const uint64_t left = c.desc.length - lo;
    const uint32_t n = uint32_t(left < uint64_t(TILE) ? left : uint64_t(TILE));
    rec.src = c.desc.src + lo * src_frame_bytes;
    rec.dst = c.desc.dst + lo * dst_frame_bytes;
    rec.frames = n;
    rec.thr2 = c.desc.thr2;
    rec.pad = 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(c.desc.src) | reinterpret_cast<uintptr_t>(c.desc.dst)) & 15) == 0;
    // whole 16-byte units on both sides: groups of 4 frames when one side has 4-byte frames
    const uint32_t unit = (src_frame_bytes == 4 || dst_frame_bytes == 4) ? 4u : 2u;
    rec.bulk_frames = aligned ? (n / unit) * unit : 0u;
}

// tile_start for the CTA's own use: 256 threads sum a contiguous run of blocks each, the run
// totals are scanned across the CTA (shuffles within a warp, shared memory across warps), and
// every thread writes its run's exclusive prefix.  ts has nblocks + 1 entries in shared memory.
template <int TILE>
__device__ __forceinline__ void batch_local_scan(const BatchBulkArgs &a, unsigned long long *ts)
{
    __shared__ unsigned long long warp_total[8];
    const uint32_t per = (a.nblocks + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per, hi = lo + per < a.nblocks ? lo + per : a.nblocks;
    unsigned long long sum = 0;
    for (uint32_t b = lo; b < hi; b++)
        sum += (a.blocks[b].length + TILE - 1) / TILE;
    unsigned long long incl = sum; // inclusive scan over the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, off);
        if ((threadIdx.x & 31) >= uint32_t(off))
            incl += up;
    }
    if ((threadIdx.x & 31) == 31)
        warp_total[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned long long before = 0;
    for (uint32_t w = 0; w < (threadIdx.x >> 5); w++)
        before += warp_total[w];
    unsigned long long acc = before + incl - sum;
    for (uint32_t b = lo; b < hi; b++) {
        ts[b] = acc;
        acc += (a.blocks[b].length + TILE - 1) / TILE;
    }
    if (threadIdx.x == blockDim.x - 1)
        ts[a.nblocks] = before + incl;
    __syncthreads();
}

template <class Op, int TILE, int STAGES, bool LOCAL_SCAN>
__global__ void __launch_bounds__(256) bulk_batch_kernel(const BatchBulkArgs a)
{
    constexpr int SFB = Op::kSrcWords * 4, DFB = Op::kDstWords * 4;
    constexpr size_t IN_STAGE = size_t(TILE) * SFB, OUT_STAGE = size_t(TILE) * DFB;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *in_buf = smem;
    unsigned char *out_buf = smem + size_t(STAGES) * IN_STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(out_buf + size_t(STAGES) * OUT_STAGE);
    TileRecord *recs = reinterpret_cast<TileRecord *>(full + STAGES);
    const unsigned long long *ts = a.tile_start;
    if constexpr (LOCAL_SCAN) {
        unsigned long long *local = reinterpret_cast<unsigned long long *>(recs + STAGES);
        batch_local_scan<TILE>(a, local);
        ts = local;
    }

    // Each CTA owns one contiguous range of tiles: consecutive tiles mostly belong to the same
    // block, so the producer touches the descriptor list only at block boundaries.
    const uint64_t ntiles = ts[a.nblocks];
    const uint64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
    const uint64_t first = uint64_t(blockIdx.x) * per;
    if (first >= ntiles)
        return;
    const uint64_t mine = ntiles - first < per ? ntiles - first : per;
    const uint64_t pol = bulk::make_policy(a.load_policy), pol_store = bulk::make_policy(a.store_policy);

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++)
            bulk::mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    BatchCursor cur; // producer state, thread 0 only
    auto produce = [&](uint64_t i) { // fills recs[i % STAGES] and, for a bulk tile, starts its load
        const int s = int(i % STAGES);
        TileRecord rec;
        batch_tile_record<TILE>(a, ts, first + i, cur, rec, SFB, DFB);
        recs[s] = rec;
        if (rec.bulk_frames) {
            bulk::mbar_expect_tx(&full[s], rec.bulk_frames * SFB);
            bulk::load_g2s(in_buf + size_t(s) * IN_STAGE, rec.src, rec.bulk_frames * SFB, &full[s], pol);
        }
    };
    if (threadIdx.x == 0) {
        batch_cursor_seek(a, ts, cur, first);
        for (uint64_t i = 0; i < uint64_t(STAGES) && i < mine; i++)
            produce(i);
    }
    __syncthreads(); // the first records are visible to everybody

    uint32_t phases = 0; // bit s: parity of the next completion of full[s]; same in every thread
    for (uint64_t i = 0; i < mine; i++) {
        const int s = int(i % STAGES);
        // out_buf[s] was last read by the bulk store of tile i - STAGES (every tile commits a
        // group, empty for direct tiles, so the count of pending groups is the count of tiles).
        if (threadIdx.x == 0)
            bulk::wait_group_read<STAGES - 1>();
        __syncthreads();
        const TileRecord rec = recs[s];

        if (rec.bulk_frames) {
            bulk::mbar_wait(&full[s], (phases >> s) & 1u);
            phases ^= 1u << s;
            const unsigned char *ib = in_buf + size_t(s) * IN_STAGE;
            unsigned char *ob = out_buf + size_t(s) * OUT_STAGE;
            constexpr int FR = (Op::kSrcWords == 1 || Op::kDstWords == 1) ? 4 : 2;
            for (uint32_t v = threadIdx.x; v * FR < rec.bulk_frames; v += blockDim.x) {
                Pack<Op::kSrcWords * FR> in;
                Pack<Op::kDstWords * FR> out;
                const uint4 *ip = reinterpret_cast<const uint4 *>(ib + size_t(v) * FR * SFB);
#pragma unroll
                for (int q = 0; q < Op::kSrcWords * FR / 4; q++) {
                    uint4 t = ip[q];
                    in.w[4 * q] = t.x, in.w[4 * q + 1] = t.y, in.w[4 * q + 2] = t.z, in.w[4 * q + 3] = t.w;
                }
                Op::template apply<FR>(in, out, rec.thr2);
                uint4 *op = reinterpret_cast<uint4 *>(ob + size_t(v) * FR * DFB);
#pragma unroll
                for (int q = 0; q < Op::kDstWords * FR / 4; q++)
                    op[q] = make_uint4(out.w[4 * q], out.w[4 * q + 1], out.w[4 * q + 2], out.w[4 * q + 3]);
            }
        }
        if (rec.bulk_frames < rec.frames) { // a direct tile, or the frames past the last 16-byte unit
            BlockDesc d;
            d.src = rec.src;
            d.dst = rec.dst;
            d.length = rec.frames;
            d.thr2 = rec.thr2;
            d.reserved = 0;
            convert_span<Op>(d, rec.bulk_frames, rec.frames, threadIdx.x, blockDim.x);
        }
        bulk::fence_async_smem(); // generic-proxy writes to out_buf visible to the bulk store
        __syncthreads();          // ... and everybody is done with recs[s] and in_buf[s]

        if (threadIdx.x == 0) {
            if (rec.bulk_frames)
                bulk::store_s2g(rec.dst, out_buf + size_t(s) * OUT_STAGE, rec.bulk_frames * DFB, pol_store);
            bulk::commit_group();
            if (i + STAGES < mine)
                produce(i + STAGES);
        }
    }
    if (threadIdx.x == 0)
        bulk::wait_group_all();
}

// tile_start for a descriptor list that lives on the device: one CTA, every thread sums a
// contiguous run of blocks, the run totals are scanned in shared memory, every thread then
// writes its run's prefix.  (Host-resident lists are summed on the host.)
template <int TILE>
__global__ void __launch_bounds__(1024) batch_tile_scan_kernel(const BlockDesc *__restrict__ blocks, uint32_t nblocks,
                                                               unsigned long long *tile_start)
{
    __shared__ unsigned long long run_total[1024];
    const uint32_t per = (nblocks + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = threadIdx.x * per, hi = lo + per < nblocks ? lo + per : nblocks;
    unsigned long long sum = 0;
    for (uint32_t b = lo; b < hi; b++)
        sum += (blocks[b].length + TILE - 1) / TILE;
    run_total[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long acc = 0;
        for (uint32_t t = 0; t < blockDim.x; t++) {
            const unsigned long long v = run_total[t];
            run_total[t] = acc;
            acc += v;
        }
        tile_start[nblocks] = acc;
    }
    __syncthreads();
    unsigned long long acc = run_total[threadIdx.x];
    for (uint32_t b = lo; b < hi; b++) {
        tile_start[b] = acc;
        acc += (blocks[b].length + TILE - 1) / TILE;
    }
}

// ---------------------------------------------------------------------------------------
// Fused repeater on the bulk-async schedule: one bulk load of I2S frames per tile, RX
// conversion shared -> shared, TX conversion shared -> shared, and two bulk stores (the CF32
// intermediate -- an API-visible buffer in the reference -- and the I2S output).  24 B/frame.
// ---------------------------------------------------------------------------------------
struct BulkLoopbackArgs {
    const char *i2s_in; // 16-byte aligned
    char *cf32;         // 16-byte aligned, may be null
    char *i2s_out;      // 16-byte aligned
    uint64_t nframes;   // even
    float thr2;
    int load_policy, store_policy;
};

template <int TILE, int STAGES>
__global__ void __launch_bounds__(256) bulk_loopback_kernel(const BulkLoopbackArgs a)
{
    constexpr size_t STAGE = size_t(TILE) * 8;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *in_buf = smem;
    unsigned char *mid_buf = smem + size_t(STAGES) * STAGE;
    unsigned char *out_buf = smem + 2 * size_t(STAGES) * STAGE;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + 3 * size_t(STAGES) * STAGE);

    const uint64_t ntiles = (a.nframes + TILE - 1) / TILE;
    const uint64_t first = blockIdx.x, stride = gridDim.x;
    if (first >= ntiles)
        return;
    const uint64_t mine = (ntiles - first + stride - 1) / stride;
    const uint64_t pol = bulk::make_policy(a.load_policy), pol_store = bulk::make_policy(a.store_policy);
    auto tile_frames = [&](uint64_t i) -> uint32_t {
        const uint64_t left = a.nframes - (first + i * stride) * TILE;
        return uint32_t(left < uint64_t(TILE) ? left : uint64_t(TILE));
    };

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++)
            bulk::mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (uint64_t i = 0; i < uint64_t(STAGES) && i < mine; i++) {
            const uint32_t nf = tile_frames(i);
            bulk::mbar_expect_tx(&full[i], nf * 8);
            bulk::load_g2s(in_buf + i * STAGE, a.i2s_in + (first + i * stride) * TILE * 8, nf * 8, &full[i], pol);
        }
    }

    for (uint64_t i = 0; i < mine; i++) {
        const int s = int(i % STAGES);
        const uint32_t nf = tile_frames(i);
        // mid_buf[s] / out_buf[s] were last read by the stores of tile i - STAGES (one group per tile)
        if (threadIdx.x == 0)
            bulk::wait_group_read<STAGES - 1>();
        __syncthreads();
        bulk::mbar_wait(&full[s], uint32_t(i / STAGES) & 1u);

        const uint4 *ip = reinterpret_cast<const uint4 *>(in_buf + size_t(s) * STAGE);
        uint4 *mp = reinterpret_cast<uint4 *>(mid_buf + size_t(s) * STAGE);
        uint4 *op = reinterpret_cast<uint4 *>(out_buf + size_t(s) * STAGE);
        for (uint32_t v = threadIdx.x; v * 2 < nf; v += blockDim.x) {
            const uint4 t = ip[v];
            Pack<4> in, mid, out;
            in.w[0] = t.x, in.w[1] = t.y, in.w[2] = t.z, in.w[3] = t.w;
            RxCf32::apply<2>(in, mid, 0.0f);
            TxCf32::apply<2>(mid, out, a.thr2);
            if (a.cf32)
                mp[v] = make_uint4(mid.w[0], mid.w[1], mid.w[2], mid.w[3]);
            op[v] = make_uint4(out.w[0], out.w[1], out.w[2], out.w[3]);
        }
        bulk::fence_async_smem();
        __syncthreads();

        if (threadIdx.x == 0) {
            const uint64_t at = (first + i * stride) * TILE * 8;
            if (a.cf32)
                bulk::store_s2g(a.cf32 + at, mid_buf + size_t(s) * STAGE, nf * 8, pol_store);
            bulk::store_s2g(a.i2s_out + at, out_buf + size_t(s) * STAGE, nf * 8, pol_store);
            bulk::commit_group();
            const uint64_t nxt = i + STAGES;
            if (nxt < mine) {
                const uint32_t nnf = tile_frames(nxt);
                bulk::mbar_expect_tx(&full[s], nnf * 8);
                bulk::load_g2s(in_buf + size_t(s) * STAGE, a.i2s_in + (first + nxt * stride) * TILE * 8, nnf * 8,
                               &full[s], pol);
            }
        }
    }
    if (threadIdx.x == 0)
        bulk::wait_group_all();
}

// ---------------------------------------------------------------------------------------
// Small synchronous calls on host buffers (readStream / writeStream of one period,
// SoapySX.cpp:451: 256 frames = 2 KiB): the kernel reads and writes the pinned host buffers
// across PCIe itself and, when the last CTA is done, stores the call's sequence number into a
// flag in pinned host memory.  The host spins on that flag instead of synchronising the stream,
// which takes the driver's completion detection (several microseconds) out of a call whose whole
// duration is of that order.
// ---------------------------------------------------------------------------------------
struct FlaggedArgs {
    BlockDesc block;              // src / dst are device-visible addresses of host (or device) memory
    unsigned int *arrivals;       // device memory, zero between launches
    unsigned long long *h_flag;   // pinned host memory
    unsigned long long seq;
};

template <class Op>
__global__ void __launch_bounds__(256) flagged_convert_kernel(const FlaggedArgs a)
{
    convert_span<Op>(a.block, 0, a.block.length, blockIdx.x * blockDim.x + threadIdx.x,
                     gridDim.x * blockDim.x);
    __threadfence_system(); // this thread's stores are visible to the host ...
    __syncthreads();        // ... and so are every other thread's of this CTA
    if (threadIdx.x == 0) {
        const unsigned int arrived = atomicAdd(a.arrivals, 1u) + 1u;
        if (arrived == gridDim.x) {
            *a.arrivals = 0; // ready for the next launch (launches on one stream do not overlap)
            __threadfence_system();
            asm volatile("st.global.wt.u64 [%0], %1;" ::"l"(a.h_flag), "l"(a.seq) : "memory");
        }
    }
}

// ---------------------------------------------------------------------------------------
// Silence, synthetic capture frames, statistics
// ---------------------------------------------------------------------------------------
__global__ void fill_silence_kernel(char *i2s, uint64_t nframes)
{
    // nframes may start on an 8-byte boundary only: frame-wide stores.
    Pack<2> z;
    z.w[0] = z.w[1] = 0;
    for (uint64_t f = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; f < nframes;
         f += uint64_t(gridDim.x) * blockDim.x)
        st_stream<8>(i2s + f * 8, z);
}

__global__ void synth_frames_kernel(char *i2s, uint64_t first_frame, uint64_t nframes, uint64_t seed)
{
    for (uint64_t f = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; f < nframes;
         f += uint64_t(gridDim.x) * blockDim.x) {
        uint64_t z = sx_synth_frame(seed, first_frame + f);
        Pack<2> p;
        p.w[0] = uint32_t(z);
        p.w[1] = uint32_t(z >> 32);
        st_stream<8>(i2s + f * 8, p);
    }
}

struct StatsAcc {
    unsigned long long sum, wsum, x, count, tx_on, rail;
};

__device__ __forceinline__ void stats_word(StatsAcc &s, uint32_t w, uint64_t idx)
{
    s.sum += w;
    s.wsum += uint64_t(w) * (2 * idx + 1);
    s.x ^= w;
    s.tx_on += (!(idx & 1) && (w & 2u)) ? 1 : 0;
    uint32_t top = w & 0xFFFFFFFCu;
    s.rail += (top == 0x7FFFFFFCu || top == 0x80000000u) ? 1 : 0;
}

// words must be 4-byte aligned; base_index is the global index of words[0].  The body runs
// on 128-bit loads; up to three leading and trailing words are taken one at a time.
__global__ void stats_kernel(const uint32_t *__restrict__ words, uint64_t nwords,
                             uint64_t base_index, StatsAcc *acc)
{
    StatsAcc s = {0, 0, 0, 0, 0, 0};
    const uint64_t tid = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint64_t nthreads = uint64_t(gridDim.x) * blockDim.x;

    uint64_t head = ((16 - (reinterpret_cast<uintptr_t>(words) & 15)) & 15) / 4;
    if (head > nwords)
        head = nwords;
    const uint64_t nvec = (nwords - head) / 4;
    const uint4 *vec = reinterpret_cast<const uint4 *>(words + head);
    for (uint64_t v = tid; v < nvec; v += nthreads) {
        uint4 q = vec[v];
        uint64_t i = base_index + head + 4 * v;
        stats_word(s, q.x, i);
        stats_word(s, q.y, i + 1);
        stats_word(s, q.z, i + 2);
        stats_word(s, q.w, i + 3);
    }
    const uint64_t tail_start = head + 4 * nvec;
    const uint64_t nedge = head + (nwords - tail_start);
    if (tid < nedge) {
        uint64_t i = tid < head ? tid : tail_start + (tid - head);
        stats_word(s, words[i], base_index + i);
    }

#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        s.sum += __shfl_xor_sync(0xffffffffu, s.sum, off);
        s.wsum += __shfl_xor_sync(0xffffffffu, s.wsum, off);
        s.x ^= __shfl_xor_sync(0xffffffffu, s.x, off);
        s.tx_on += __shfl_xor_sync(0xffffffffu, s.tx_on, off);
        s.rail += __shfl_xor_sync(0xffffffffu, s.rail, off);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&acc->sum, s.sum);
        atomicAdd(&acc->wsum, s.wsum);
        atomicXor(&acc->x, s.x);
        atomicAdd(&acc->tx_on, s.tx_on);
        atomicAdd(&acc->rail, s.rail);
    }
}

} // namespace sx
