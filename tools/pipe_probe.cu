// pipe_probe.cu -- do fp64, ALU and FMA-pipe instructions of one warp scheduler overlap on sm_100a?
// Each pattern runs independent dependency chains per thread, unrolled; reports cycles per loop
// iteration per warp for 1, 2 and 4 warps per scheduler.  Build: nvcc -arch=sm_100a -O3 -o pipe_probe
#include <cstdio>
#include <cuda_runtime.h>

#define CH 8
template <int MODE>
__global__ void probe(double *out, int iters, long long *cyc) {
    double d[CH];
    int a[CH];
    float f[CH];
    for (int k = 0; k < CH; ++k) { d[k] = 1.0 + threadIdx.x * 1e-9 + k; a[k] = threadIdx.x + k; f[k] = 1.0f + k; }
    const double m = 1.0000001, c = 1e-12;
    const int lim = 1 << 30;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < CH; ++k) {
            if (MODE == 0 || (MODE >= 2 && MODE <= 6)) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(m), "d"(c));
            if (MODE == 1 || MODE == 2) asm volatile("min.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(lim - it));        // ALU
            if (MODE == 3) asm volatile("mad.lo.s32 %0, %0, 3, %1;" : "+r"(a[k]) : "r"(it));                       // FMA pipe (IMAD)
            if (MODE == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(1.0001f), "f"(1e-6f));     // FFMA
            if (MODE == 5) { asm volatile("min.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(lim - it));
                             asm volatile("min.s32 %0, %0, %1;" : "+r"(a[k]) : "r"(lim - 2 * it)); }               // fp64 + 2 ALU
            if (MODE == 7) asm volatile("{.reg .pred p; setp.gt.s32 p, %1, 5; selp.b32 %0, %0, %1, p;}" : "+r"(a[k]) : "r"(it));   // ISETP + SEL
            if (MODE == 8) asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; @p add.s32 %2, %2, 1;}" :: "d"(d[k]), "d"(m), "r"(a[k]));  // DSETP only
            if (MODE == 9) {  // the M-state sum of the pair-HMM kernel with the e^-10 rule, one cell
                double x = d[k], y = d[(k + 1) % CH] * 1e-5, z = d[(k + 2) % CH] * 2e-5, r;
                const double s3 = x + y + z;
                asm volatile("{\n\t.reg .pred p, q, pa, pb, pc, t, u;\n\t.reg .f64 hi, mx, thr;\n\t"
                    "setp.gt.f64 p, %1, %2;\n\tselp.f64 hi, %1, %2, p;\n\tsetp.gt.f64 q, hi, %3;\n\t"
                    "selp.f64 mx, hi, %3, q;\n\tmul.f64 thr, mx, 0d3F07CD79B5647C9B;\n\t"
                    "setp.ge.f64 pa, %1, thr;\n\tsetp.ge.f64 pb, %2, thr;\n\tsetp.ge.f64 pc, %3, thr;\n\t"
                    "and.pred t, pa, pb;\n\tor.pred u, pa, pb;\n\tand.pred u, u, pc;\n\tor.pred t, t, u;\n\t"
                    "selp.f64 %0, %4, mx, t;\n\t}" : "=d"(r) : "d"(x), "d"(y), "d"(z), "d"(s3));
                d[k] = r * 0.999;
            }
            if (MODE == 6) { asm volatile("{.reg .pred p; setp.gt.f64 p, %0, %1; selp.f64 %0, %0, %1, p;}" : "+d"(d[k]) : "d"(m)); }  // + DSETP + 2 FSEL
        }
    }
    long long t1 = clock64();
    double s = 0; int sa = 0; float sf = 0;
    for (int k = 0; k < CH; ++k) { s += d[k]; sa += a[k]; sf += f[k]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + sa + sf;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char *name, double *d_out, long long *d_cyc) {
    const int iters = 4096;
    printf("%-34s", name);
    for (int threads : {128, 256, 512, 1024}) {
        probe<MODE><<<148, threads>>>(d_out, iters, d_cyc);
        probe<MODE><<<148, threads>>>(d_out, iters, d_cyc);
        cudaDeviceSynchronize();
        long long c;
        cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
        printf("  %2d w/sched: %6.2f cyc/iter/chain", threads / 128, (double)c / iters / CH);
    }
    printf("\n");
}

int main() {
    double *d_out; long long *d_cyc;
    cudaMalloc(&d_out, 148 * 1024 * 8); cudaMalloc(&d_cyc, 8);
    run<0>("DFMA", d_out, d_cyc);
    run<1>("IMNMX (ALU)", d_out, d_cyc);
    run<2>("DFMA + IMNMX", d_out, d_cyc);
    run<3>("DFMA + IMAD", d_out, d_cyc);
    run<4>("DFMA + FFMA", d_out, d_cyc);
    run<5>("DFMA + 2 IMNMX", d_out, d_cyc);
    run<6>("DFMA + DSETP + 64-bit select", d_out, d_cyc);
    run<7>("ISETP + SEL", d_out, d_cyc);
    run<8>("DSETP", d_out, d_cyc);
    run<9>("sum3 approx cell (5 DSETP 4 DMUL 2 DADD 6 FSEL)", d_out, d_cyc);
    return 0;
}
