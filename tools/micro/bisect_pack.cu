// Micro-benchmark: the reference's 24-step cubic bisection (make_intersection_1.comp:392-436), scalar fp32
// (21 instructions per step) against packed f32x2 forms, on sm_100a. Prints throughput and whether the
// packed forms give bit-identical results. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t f2u(float f) { return __float_as_uint(f); }
__device__ __forceinline__ uint64_t pk(float a0, float a1) { uint64_t a; asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1)); return a; }
__device__ __forceinline__ void upk(uint64_t r, float &r0, float &r1) { asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r)); }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2z(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float lerpf(float a, float b, float t) { return __fadd_rn(a, __fmul_rn(t, __fsub_rn(b, a))); }

#define SELECT(t0, t1, s_last, neg0, tm)                                                                     \
    asm("{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %3, 0;\n\tsetp.lt.xor.s32 p, %2, 0, q;\n\t"              \
        "selp.f32 %0, %0, %4, p;\n\tselp.f32 %1, %4, %1, p;\n\t}"                                          \
        : "+f"(t0), "+f"(t1) : "r"(s_last), "r"((int)neg0), "f"(tm))

template <int VAR>
__device__ __forceinline__ float solve(float c0, float c1, float c2, float c3, float t0, float t1, float cst) {
    const float d01 = __fsub_rn(c1, c0), d12 = __fsub_rn(c2, c1), d23 = __fsub_rn(c3, c2);
    float vt0;
    {
        const float a0 = __fadd_rn(c0, __fmul_rn(t0, d01)), a1 = __fadd_rn(c1, __fmul_rn(t0, d12)), a2 = __fadd_rn(c2, __fmul_rn(t0, d23));
        vt0 = lerpf(lerpf(a0, a1, t0), lerpf(a1, a2, t0), t0);
    }
    float t_solve = t0;
    if (vt0 == cst) return t_solve;
    const float raw_t0 = t0;
    const bool neg0 = (int)f2u(__fsub_rn(vt0, cst)) < 0;
    uint32_t s_last = 0;
    const uint64_t D01 = pk(d01, d12), D12 = pk(d12, d23), C01 = pk(c0, c1), C12 = pk(c1, c2);
#pragma unroll 4
    for (int j = 0; j < 24; ++j) {
        const float tm = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
        float vtm;
        if (VAR == 0) {
            const float a0 = __fadd_rn(c0, __fmul_rn(tm, d01)), a1 = __fadd_rn(c1, __fmul_rn(tm, d12)), a2 = __fadd_rn(c2, __fmul_rn(tm, d23));
            vtm = lerpf(lerpf(a0, a1, tm), lerpf(a1, a2, tm), tm);
        } else if (VAR == 1) {  // fully packed, .ftz on the products keeps ptxas from contracting them into FFMA2
            const uint64_t T = pk(tm, tm);
            const uint64_t A01 = add2(C01, mul2z(T, D01)), A12 = add2(C12, mul2z(T, D12));
            float b0, b1;
            upk(add2(A01, mul2z(T, sub2(A12, A01))), b0, b1);
            vtm = lerpf(b0, b1, tm);
        } else if (VAR == 2) {  // packed products, scalar sums
            const uint64_t T = pk(tm, tm);
            float m0, m1; upk(mul2(T, D01), m0, m1);
            const float a0 = __fadd_rn(c0, m0), a1 = __fadd_rn(c1, m1), a2 = __fadd_rn(c2, __fmul_rn(tm, d23));
            upk(mul2(T, sub2(pk(a1, a2), pk(a0, a1))), m0, m1);
            vtm = lerpf(__fadd_rn(a0, m0), __fadd_rn(a1, m1), tm);
        } else {  // contracted (FFMA2): NOT the reference arithmetic, speed comparison only
            const uint64_t T = pk(tm, tm);
            const uint64_t A01 = add2(C01, mul2(T, D01)), A12 = add2(C12, mul2(T, D12));
            float b0, b1;
            upk(add2(A01, mul2(T, sub2(A12, A01))), b0, b1);
            vtm = lerpf(b0, b1, tm);
        }
        t_solve = tm;
        s_last = f2u(__fsub_rn(vtm, cst));
        SELECT(t0, t1, s_last, neg0, tm);
    }
    if (fabsf(__uint_as_float(s_last)) > 1.f) t_solve = raw_t0;
    return t_solve;
}

template <int VAR>
__global__ void __launch_bounds__(128) k(const float4 *__restrict__ cs, const float *__restrict__ line, float *__restrict__ out, int n, int reps) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = cs[i];
    float t = 0.f, acc = 0.f;
    const float l0 = line[i];
    for (int r = 0; r < reps; ++r) {  // a chain of crossings, like the walk along a piece
        t = solve<VAR>(c.x, c.y, c.z, c.w, t, 1.0f, l0 + 2.0f * r);
        acc = __uint_as_float(f2u(acc) ^ f2u(t)) ;
    }
    out[i] = acc;
}

int main() {
    const int n = 148 * 8 * 128 * 4, reps = 16;
    std::vector<float4> cs(n); std::vector<float> line(n);
    uint64_t s = 0x5CA71E01ull;
    auto rnd = [&]() { s += 0x9E3779B97F4A7C15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31; return (float)((z >> 40) * (1.0 / 16777216.0)); };
    for (int i = 0; i < n; ++i) {  // monotone increasing cubics spanning ~40 px
        float a = rnd() * 3800.f, b = a + rnd() * 15.f, c = b + rnd() * 15.f, d = c + 2.f + rnd() * 15.f;
        cs[i] = make_float4(a, b, c, d);
        line[i] = 2.0f * (float)(int)(a * 0.5f) + 2.0f;
    }
    float4 *dcs; float *dl, *dout[4];
    cudaMalloc(&dcs, n * sizeof(float4)); cudaMalloc(&dl, n * 4);
    for (auto &p : dout) cudaMalloc(&p, n * 4);
    cudaMemcpy(dcs, cs.data(), n * sizeof(float4), cudaMemcpyHostToDevice);
    cudaMemcpy(dl, line.data(), n * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *names[4] = {"scalar (reference arithmetic)", "packed f32x2, ftz products", "packed products + scalar sums", "contracted FFMA2 (not exact)"};
    std::vector<float> h0(n), h(n);
    for (int v = 0; v < 4; ++v) {
        float best = 1e9f;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(e0);
            if (v == 0) k<0><<<n / 128, 128>>>(dcs, dl, dout[0], n, reps);
            if (v == 1) k<1><<<n / 128, 128>>>(dcs, dl, dout[1], n, reps);
            if (v == 2) k<2><<<n / 128, 128>>>(dcs, dl, dout[2], n, reps);
            if (v == 3) k<3><<<n / 128, 128>>>(dcs, dl, dout[3], n, reps);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (it > 0 && ms < best) best = ms;
        }
        cudaMemcpy(v == 0 ? h0.data() : h.data(), dout[v], n * 4, cudaMemcpyDeviceToHost);
        long diff = 0;
        if (v) for (int i = 0; i < n; ++i) diff += (reinterpret_cast<uint32_t &>(h[i]) != reinterpret_cast<uint32_t &>(h0[i]));
        printf("%-34s %.4f ms  %.1f G bisection steps/s  differing results vs scalar: %ld of %d  (err %s)\n", names[v], best,
               (double)n * reps * 24 / best / 1e6, diff, n, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
