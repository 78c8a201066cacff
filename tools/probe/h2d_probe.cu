// H2D of one 1242x375 image from page-locked memory: contiguous vs pitched (2-D) copy, device time (CUDA events).
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
    const int W = 1242, H = 375, P = 1280;
    unsigned char *h, *d;
    cudaMallocHost(&h, (size_t)P * H);
    cudaMalloc(&d, (size_t)P * H * 2);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int mode = 0; mode < 3; mode++)
    {
        float best = 1e9f, sum = 0;
        for (int it = 0; it < 60; it++)
        {
            cudaEventRecord(a, s);
            if (mode == 0)
                cudaMemcpyAsync(d, h, (size_t)W * H, cudaMemcpyHostToDevice, s);
            else if (mode == 1)
                cudaMemcpy2DAsync(d, P, h, W, W, H, cudaMemcpyHostToDevice, s);
            else
                cudaMemcpyAsync(d, h, (size_t)P * H, cudaMemcpyHostToDevice, s);
            cudaEventRecord(b, s);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (it >= 10)
                sum += ms, best = ms < best ? ms : best;
        }
        printf("%-40s best %.1f us  mean %.1f us\n", mode == 0 ? "1-D packed (465750 B)" : mode == 1 ? "2-D 1242 -> pitch 1280" : "1-D pitched (480000 B)",
               1e3f * best, 1e3f * sum / 50);
    }
    return 0;
}
