// Block-row sharding of ONE image over the GPUs of a box: the only data that has to cross GPUs on the way to the encoded
// chain is the last distributed fp32 mip level (a few hundred KB), which band 0 needs to build the small tail levels from.
// Every band stores its rows of that level straight into band 0's memory (peer stores over NVLink / NVSwitch) and raises
// a flag with release semantics; band 0 waits for the flags with acquire loads on a second stream while its own encodes run.
// Nothing in the reference corresponds to this (it has no multi-device path); the data it moves is FloatImage rows
// (src/nvimage/FloatImage.h:193-229 layout).
#pragma once
#include "../nvb_common.cuh"

namespace nvb {

// exchange buffer: 64 arrival flags, the acknowledge word, then two copies (sequence parity) of the level
#define NVB_XCHG_FLAGS 64
#define NVB_XCHG_ACK 64       // index of the acknowledge word (unsigned units)
#define NVB_XCHG_HEADER 512   // bytes
#define NVB_XCHG_TIMEOUT_NS 10000000000ull

struct ExportRowsParams {
    const float *src;  // this band's rows of the level: [4][hl][w], chunks concatenated
    float *dst;        // the whole level: [4][h][w] (possibly peer memory)
    int w, hl, h;
    int chunk_rows;    // rows of one chunk at this level
    int n, b;          // bands, this band
};

__global__ void __launch_bounds__(256) k_export_rows(ExportRowsParams P) {
    const size_t per_plane = (size_t)P.w * P.hl, total = per_plane * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i / per_plane);
        const size_t r = i - (size_t)c * per_plane;
        const int y = (int)(r / P.w), x = (int)(r - (size_t)y * P.w);
        const int j = y / P.chunk_rows, ry = y - j * P.chunk_rows;
        const int gy = (j * P.n + P.b) * P.chunk_rows + ry;
        P.dst[(size_t)c * P.w * P.h + (size_t)gy * P.w + x] = P.src[i];
    }
}

NVB_DEV unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
NVB_DEV void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
NVB_DEV unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// after k_export_rows on the same stream: the rows are ordered before the flag (kernel boundary, then a system-scope release)
__global__ void k_xchg_signal(unsigned *word, unsigned seq) {
    __threadfence_system();
    st_release_sys(word, seq);
}

// wait until words[0..n) have all reached `seq` (wrap-safe compare); *fault is host-mapped and set when the wait times out
__global__ void k_xchg_wait(const unsigned *words, int n, unsigned seq, unsigned *fault) {
    const int t = threadIdx.x;
    if (t < n) {
        const unsigned long long t0 = global_ns();
        while ((int)(ld_acquire_sys(words + t) - seq) < 0) {
            if (global_ns() - t0 > NVB_XCHG_TIMEOUT_NS) {
                *fault = 1u;
                break;
            }
            __nanosleep(500);
        }
    }
    __syncthreads();
    __threadfence_system();
}

}  // namespace nvb
