// Result wire format on the GPU.
//
// The reference worker answers a WorkOrder with the protobuf message
//     message TraceResults { message Colour { uint32 r = 1; uint32 g = 2; uint32 b = 3; }  repeated Colour results = 1; }
// (/root/reference/shared/comms/comms.proto:34-41), built one heap-allocated Colour per pixel at
// worker/distributed/main.go:52-54,79-86 and marshalled by grpc-go.  At 4K that is 8.3 M sub-messages per frame and, per
// the survey, what caps the end-to-end frame rate -- not rendering.  These kernels turn the region's RGB bytes (the order
// rt_render_region produces, i*h+j) into exactly the bytes proto.Marshal emits for that message, so that a GPU worker can
// hand grpc a ready-made payload and an unmodified master decodes it.
//
// proto3 wire form of one pixel:  0x0A  len  [0x08 varint(r)] [0x10 varint(g)] [0x18 varint(b)]   -- a field whose value is 0
// is omitted (proto3 default), a value < 128 is a 1-byte varint, 128..255 takes 2 bytes; len <= 9, so 2..11 bytes a pixel.
//
// Round-2 experiment, opt-in (RT_B200_WIRE_ONEPASS=1): ONE pass.  Measured on B200 on the 4K C3 frame: 72.7 us against 57.3 us for
// the three kernels below -- the look-back chain over 8100 blocks costs more latency than the second read of 25 MB it saves.
//   k_wire_onepass    per 1024-pixel block: sizes, block-local scan, decoupled look-back over the preceding blocks' aggregates
//                     (Merrill & Garland's single-pass scan: a block publishes its aggregate, then its inclusive prefix, in one
//                     64-bit status word), encode into shared memory at the output's 16-byte phase, copy out with uint4 stores.
//                     The pixels are read once (3 B), 2..11 B are written, the last block writes the message length.
// Default:
//   W1 k_wire_sizes   bytes of every 1024-pixel block                      (reads 3 B / pixel)
//   W2 k_wire_scan    exclusive scan of the block sizes + total length     (one CTA)
//   W3 k_wire_encode  block-local scan, encode into shared memory, copy out with 16-byte stores  (3 B read, 2..11 B written)
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#include "rt_internal.h"

namespace {

constexpr int WIRE_THREADS = 256;
constexpr int WIRE_PIX_PER_THREAD = 4;
constexpr int WIRE_BLOCK_PIX = WIRE_THREADS * WIRE_PIX_PER_THREAD;     // 1024 pixels = 3072 B in, <= 11264 B out

__device__ __forceinline__ uint32_t chan_bytes(uint32_t v) { return v == 0u ? 0u : (v < 128u ? 2u : 3u); }
__device__ __forceinline__ uint32_t pixel_bytes(uint32_t r, uint32_t g, uint32_t b) { return 2u + chan_bytes(r) + chan_bytes(g) + chan_bytes(b); }

// the 4 pixels (12 bytes) of this thread; pixels beyond n read as absent (size 0)
__device__ __forceinline__ void load_pixels(const uint8_t *__restrict__ rgb, uint64_t first, uint64_t n, uint32_t px[WIRE_PIX_PER_THREAD][3], int &count) {
    count = first >= n ? 0 : (int)(n - first < (uint64_t)WIRE_PIX_PER_THREAD ? n - first : (uint64_t)WIRE_PIX_PER_THREAD);
    if (count == WIRE_PIX_PER_THREAD && ((uintptr_t)(rgb + first * 3) & 3u) == 0) {
        const uint32_t *p = reinterpret_cast<const uint32_t *>(rgb + first * 3);
        uint32_t w[3] = { __ldg(p), __ldg(p + 1), __ldg(p + 2) };
#pragma unroll
        for (int k = 0; k < 12; k++) px[k / 3][k % 3] = (w[k >> 2] >> (8 * (k & 3))) & 0xffu;
    } else {
#pragma unroll
        for (int q = 0; q < WIRE_PIX_PER_THREAD; q++)
#pragma unroll
            for (int c = 0; c < 3; c++) px[q][c] = q < count ? (uint32_t)__ldg(rgb + (first + q) * 3 + c) : 0u;
    }
}

__device__ __forceinline__ uint32_t thread_bytes(const uint32_t px[WIRE_PIX_PER_THREAD][3], int count) {
    uint32_t s = 0;
#pragma unroll
    for (int q = 0; q < WIRE_PIX_PER_THREAD; q++) s += q < count ? pixel_bytes(px[q][0], px[q][1], px[q][2]) : 0u;
    return s;
}

// block-wide exclusive scan of one value per thread (256 threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t *total) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= (unsigned)d) inc += o;
    }
    if (lane == 31u) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (unsigned)(WIRE_THREADS / 32) ? s_warp[lane] : 0u;
#pragma unroll
        for (int d = 1; d < WIRE_THREADS / 32; d <<= 1) {
            uint32_t o = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= (unsigned)d) w += o;
        }
        if (lane < (unsigned)(WIRE_THREADS / 32)) s_warp[lane] = w;          // inclusive over warps
    }
    __syncthreads();
    *total = s_warp[WIRE_THREADS / 32 - 1];
    return inc - v + (warp ? s_warp[warp - 1] : 0u);
}

__global__ void __launch_bounds__(WIRE_THREADS) k_wire_sizes(const uint8_t *__restrict__ rgb, uint64_t n, uint32_t *__restrict__ block_bytes) {
    __shared__ uint32_t s_warp[WIRE_THREADS / 32];
    uint32_t px[WIRE_PIX_PER_THREAD][3];
    int count;
    load_pixels(rgb, (uint64_t)blockIdx.x * WIRE_BLOCK_PIX + (uint64_t)threadIdx.x * WIRE_PIX_PER_THREAD, n, px, count);
    uint32_t v = thread_bytes(px, count);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31u) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
#pragma unroll
        for (int w = 0; w < WIRE_THREADS / 32; w++) t += s_warp[w];
        block_bytes[blockIdx.x] = t;
    }
}

// one CTA: block_offset[b] = sum of block_bytes[0..b), *total = the message length
__global__ void __launch_bounds__(1024) k_wire_scan(const uint32_t *__restrict__ block_bytes, uint32_t n_blocks, uint64_t *__restrict__ block_offset,
                                                     uint64_t *__restrict__ total) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += 1024u) {
        const uint32_t b = base + threadIdx.x;
        const uint64_t v = b < n_blocks ? (uint64_t)block_bytes[b] : 0ull;
        uint64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= (unsigned)d) inc += o;
        }
        if (lane == 31u) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint64_t w = s_warp[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint64_t o = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= (unsigned)d) w += o;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const uint64_t carry = s_carry;
        if (b < n_blocks) block_offset[b] = carry + inc - v + (warp ? s_warp[warp - 1] : 0ull);
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}

__device__ __forceinline__ uint32_t put_channel(uint8_t *p, uint32_t tag, uint32_t v) {
    if (v == 0u) return 0u;
    p[0] = (uint8_t)tag;
    if (v < 128u) { p[1] = (uint8_t)v; return 2u; }
    p[1] = (uint8_t)(v | 0x80u);              // low 7 bits + continuation
    p[2] = (uint8_t)(v >> 7);
    return 3u;
}

__global__ void __launch_bounds__(WIRE_THREADS) k_wire_encode(const uint8_t *__restrict__ rgb, uint64_t n, const uint64_t *__restrict__ block_offset,
                                                               uint8_t *__restrict__ out, uint64_t capacity) {
    // the block's bytes are staged at s_bytes[pad ...] with pad = (global offset mod 16), so that shared-memory chunk k and
    // global chunk k have the same 16-byte alignment and whole chunks move as uint4
    __shared__ __align__(16) uint8_t s_bytes[16 + WIRE_BLOCK_PIX * 11 + 16];
    __shared__ uint32_t s_warp[WIRE_THREADS / 32];
    uint32_t px[WIRE_PIX_PER_THREAD][3];
    int count;
    load_pixels(rgb, (uint64_t)blockIdx.x * WIRE_BLOCK_PIX + (uint64_t)threadIdx.x * WIRE_PIX_PER_THREAD, n, px, count);
    uint32_t total;
    const uint32_t off = block_exclusive_scan(thread_bytes(px, count), s_warp, &total);
    const uint64_t g0 = block_offset[blockIdx.x];
    if (g0 + total > capacity) return;                       // caller compares the reported length with its capacity
    const uint32_t pad = (uint32_t)((uintptr_t)(out + g0) & 15u);
    uint8_t *p = s_bytes + pad + off;
#pragma unroll
    for (int q = 0; q < WIRE_PIX_PER_THREAD; q++) {
        if (q < count) {
            uint32_t len = 0;
            len += put_channel(p + 2 + len, 0x08u, px[q][0]);
            len += put_channel(p + 2 + len, 0x10u, px[q][1]);
            len += put_channel(p + 2 + len, 0x18u, px[q][2]);
            p[0] = 0x0Au;                                    // field 1, wire type 2 (length-delimited)
            p[1] = (uint8_t)len;                             // <= 9
            p += 2 + len;
        }
    }
    __syncthreads();
    uint8_t *gbase = out + g0 - pad;                         // 16-byte aligned
    const uint32_t end = pad + total;                        // valid bytes are s_bytes[pad, end)
    const uint32_t n_chunks = (end + 15u) / 16u;
    for (uint32_t c = threadIdx.x; c < n_chunks; c += WIRE_THREADS) {
        const uint32_t lo = c * 16u, hi = lo + 16u;
        if (lo >= pad && hi <= end) {
            *reinterpret_cast<uint4 *>(gbase + lo) = *reinterpret_cast<const uint4 *>(s_bytes + lo);
        } else {                                             // first / last chunk shared with the neighbouring blocks: own bytes only
            for (uint32_t k = (lo > pad ? lo : pad); k < (hi < end ? hi : end); k++) gbase[k] = s_bytes[k];
        }
    }
}

// ---- single pass ---------------------------------------------------------------------------------------------------------
// status[b]: bits 63..62 = 0 nothing yet, 1 = aggregate of block b, 2 = inclusive prefix up to and including block b; low 62 bits = bytes
constexpr unsigned long long WIRE_FLAG_AGG = 1ull << 62, WIRE_FLAG_INC = 2ull << 62, WIRE_VAL_MASK = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(WIRE_THREADS) k_wire_onepass(const uint8_t *__restrict__ rgb, uint64_t n, unsigned long long *status,
                                                                unsigned int *ticket, uint32_t n_blocks, uint8_t *__restrict__ out,
                                                                uint64_t capacity, uint64_t *__restrict__ total_len) {
    __shared__ __align__(16) uint8_t s_bytes[16 + WIRE_BLOCK_PIX * 11 + 16];
    __shared__ uint32_t s_warp[WIRE_THREADS / 32];
    __shared__ uint32_t s_block;
    __shared__ unsigned long long s_prefix;
    // blocks take their index from a ticket in the order they START, so every predecessor a block waits for is already running
    if (threadIdx.x == 0) s_block = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t b = s_block;
    uint32_t px[WIRE_PIX_PER_THREAD][3];
    int count;
    load_pixels(rgb, (uint64_t)b * WIRE_BLOCK_PIX + (uint64_t)threadIdx.x * WIRE_PIX_PER_THREAD, n, px, count);
    uint32_t total;
    const uint32_t off = block_exclusive_scan(thread_bytes(px, count), s_warp, &total);
    // publish the aggregate, look back for the exclusive prefix (warp 0; one status word per lane and step)
    if (threadIdx.x < 32) {
        const unsigned lane = threadIdx.x;
        if (lane == 0) st_status(status + b, (b == 0 ? WIRE_FLAG_INC : WIRE_FLAG_AGG) | (unsigned long long)total);
        unsigned long long prefix = 0;
        if (b > 0) {
            int64_t look = (int64_t)b - 1 - (int64_t)lane;            // lane k inspects block b-1-k
            for (;;) {
                unsigned long long v = look >= 0 ? ld_status(status + look) : WIRE_FLAG_INC;      // before block 0: prefix 0
                // wait until every inspected block has published something
                while (__any_sync(0xffffffffu, (v >> 62) == 0ull)) v = look >= 0 ? ld_status(status + look) : WIRE_FLAG_INC;
                const unsigned inc_mask = __ballot_sync(0xffffffffu, (v >> 62) == 2ull);
                // sum the aggregates of the lanes in front of the first inclusive prefix, plus that prefix
                const int first_inc = inc_mask ? __ffs(inc_mask) - 1 : 32;
                unsigned long long part = (int)lane <= first_inc ? (v & WIRE_VAL_MASK) : 0ull;
                if ((int)lane > first_inc) part = 0ull;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                prefix += part;
                if (inc_mask) break;
                look -= 32;
            }
        }
        if (lane == 0) {
            s_prefix = prefix;
            if (b > 0) st_status(status + b, WIRE_FLAG_INC | (prefix + (unsigned long long)total));
            if (b == n_blocks - 1) *total_len = prefix + (unsigned long long)total;
        }
    }
    __syncthreads();
    const uint64_t g0 = s_prefix;
    if (g0 + total > capacity) return;                       // caller compares the reported length with its capacity
    const uint32_t pad = (uint32_t)((uintptr_t)(out + g0) & 15u);
    uint8_t *p = s_bytes + pad + off;
#pragma unroll
    for (int q = 0; q < WIRE_PIX_PER_THREAD; q++) {
        if (q < count) {
            uint32_t len = 0;
            len += put_channel(p + 2 + len, 0x08u, px[q][0]);
            len += put_channel(p + 2 + len, 0x10u, px[q][1]);
            len += put_channel(p + 2 + len, 0x18u, px[q][2]);
            p[0] = 0x0Au;
            p[1] = (uint8_t)len;
            p += 2 + len;
        }
    }
    __syncthreads();
    uint8_t *gbase = out + g0 - pad;
    const uint32_t end = pad + total;
    const uint32_t n_chunks = (end + 15u) / 16u;
    for (uint32_t c = threadIdx.x; c < n_chunks; c += WIRE_THREADS) {
        const uint32_t lo = c * 16u, hi = lo + 16u;
        if (lo >= pad && hi <= end) {
            *reinterpret_cast<uint4 *>(gbase + lo) = *reinterpret_cast<const uint4 *>(s_bytes + lo);
        } else {
            for (uint32_t k = (lo > pad ? lo : pad); k < (hi < end ? hi : end); k++) gbase[k] = s_bytes[k];
        }
    }
}

}  // namespace

size_t rt_wire_scratch_bytes(uint64_t n_pixels) {
    const uint64_t nb = (n_pixels + WIRE_BLOCK_PIX - 1) / WIRE_BLOCK_PIX;
    return (size_t)(16 + nb * 4 + 16 + nb * 8 + 16);          // both forms fit: {block_bytes, block_offset} or {status words, ticket}
}

// d_scratch: rt_wire_scratch_bytes(n) bytes, 16-byte aligned.  d_len receives the message length (also when it exceeds capacity,
// in which case nothing beyond `capacity` is written).
cudaError_t rt_launch_wire_encode(const uint8_t *d_rgb, uint64_t n_pixels, uint8_t *d_out, uint64_t capacity, uint64_t *d_len,
                                  void *d_scratch, cudaStream_t st) {
    if (n_pixels == 0) return cudaMemsetAsync(d_len, 0, sizeof(uint64_t), st);
    const uint64_t nb = (n_pixels + WIRE_BLOCK_PIX - 1) / WIRE_BLOCK_PIX;
    if (nb > 0x7fffffffull) return cudaErrorInvalidValue;
    static const bool onepass = [] { const char *v = getenv("RT_B200_WIRE_ONEPASS"); return v && *v && atoi(v) != 0; }();   // opt-in: measured SLOWER than the three kernels (72.7 vs 57.3 us per 4K frame, profiles/r02_wire_encoder_timing.json)
    if (onepass) {
        unsigned long long *status = reinterpret_cast<unsigned long long *>(d_scratch);            // nb words, then the ticket
        unsigned int *ticket = reinterpret_cast<unsigned int *>(status + nb);
        cudaError_t e = cudaMemsetAsync(d_scratch, 0, (size_t)(nb * 8 + 8), st);
        if (e != cudaSuccess) return e;
        k_wire_onepass<<<(unsigned)nb, WIRE_THREADS, 0, st>>>(d_rgb, n_pixels, status, ticket, (uint32_t)nb, d_out, capacity, d_len);
        return cudaGetLastError();
    }
    uint32_t *block_bytes = reinterpret_cast<uint32_t *>(d_scratch);
    uint64_t *block_offset = reinterpret_cast<uint64_t *>(reinterpret_cast<uint8_t *>(d_scratch) + ((nb * 4 + 15) / 16) * 16);
    k_wire_sizes<<<(unsigned)nb, WIRE_THREADS, 0, st>>>(d_rgb, n_pixels, block_bytes);
    k_wire_scan<<<1, 1024, 0, st>>>(block_bytes, (uint32_t)nb, block_offset, d_len);
    k_wire_encode<<<(unsigned)nb, WIRE_THREADS, 0, st>>>(d_rgb, n_pixels, block_offset, d_out, capacity);
    return cudaGetLastError();
}
