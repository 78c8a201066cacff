// Dependent-chain latency of fp64 operations on one thread / one warp (B200): cycles per operation.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void chain(double *out, double a, double b, int n, long long *cyc)
{
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < n; i++)
    {
#pragma unroll
        for (int k = 0; k < 16; k++)
        {
            if (OP == 0) x = fma(x, b, a);
            if (OP == 1) x = x * b + a; // -fmad=false: DMUL + DADD
            if (OP == 2) x = a / x + b;
            if (OP == 3) x = sqrt(x) + b;
            if (OP == 4) x = 1.0 / x + b;
            if (OP == 5) x = rsqrt(x) + b;
            if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1) + b;
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0)
        *cyc = t1 - t0;
}
int main()
{
    double *d;
    long long *c, h;
    cudaMalloc(&d, 8 * 1024);
    cudaMalloc(&c, 8);
    const char *names[] = {"fma", "mul+add (2 ops)", "div+add", "sqrt+add", "rcp+add", "rsqrt+add", "shfl(double)+add"};
    for (int threads : {1, 32})
        for (int op = 0; op < 7; op++)
        {
            const int n = 64;
            for (int rep = 0; rep < 2; rep++)
            {
                if (op == 0) chain<0><<<1, threads>>>(d, 1.0000001, 0.9999999, n, c);
                if (op == 1) chain<1><<<1, threads>>>(d, 1.0000001, 0.9999999, n, c);
                if (op == 2) chain<2><<<1, threads>>>(d, 1.5, 0.7, n, c);
                if (op == 3) chain<3><<<1, threads>>>(d, 1.5, 0.7, n, c);
                if (op == 4) chain<4><<<1, threads>>>(d, 1.5, 0.7, n, c);
                if (op == 5) chain<5><<<1, threads>>>(d, 1.5, 0.7, n, c);
                if (op == 6) chain<6><<<1, threads>>>(d, 1.5, 0.7, n, c);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
            printf("%2d thread(s) %-18s %.1f cycles per step\n", threads, names[op], (double)h / (n * 16));
        }
    return 0;
}
