// Which hardware warp slots (-> SM sub-partitions, slot % 4) do the warps of small blocks land in?
// usage: warp_slots <threads_per_block> <blocks_per_sm> <regs_dummy>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void k(int *out, long long spin) {
    unsigned smid, wid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if ((threadIdx.x & 31) == 0) {
        int w = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
        out[2 * w] = smid;
        out[2 * w + 1] = wid;
    }
}
int main(int argc, char **argv) {
    int tpb = argc > 1 ? atoi(argv[1]) : 64, bps = argc > 2 ? atoi(argv[2]) : 7;
    int smem = argc > 3 ? atoi(argv[3]) : 30000;
    int nblk = 148 * bps, nw = nblk * tpb / 32;
    int *d, *h = (int *)malloc(nw * 8);
    cudaMalloc(&d, nw * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<nblk, tpb, smem>>>(d, 2000000);
    cudaDeviceSynchronize();
    cudaMemcpy(h, d, nw * 8, cudaMemcpyDeviceToHost);
    // histogram of slot % 4 per SM, then summarise the per-SM (min, max) load over sub-partitions
    int hist[200][4] = {};
    for (int w = 0; w < nw; w++) hist[h[2 * w]][h[2 * w + 1] & 3]++;
    int worst = 0, best = 1 << 30;
    for (int s = 0; s < 148; s++) {
        int mx = 0, mn = 1 << 30, tot = 0;
        for (int q = 0; q < 4; q++) { mx = hist[s][q] > mx ? hist[s][q] : mx; mn = hist[s][q] < mn ? hist[s][q] : mn; tot += hist[s][q]; }
        if (s < 4) printf("SM %d: %d %d %d %d\n", s, hist[s][0], hist[s][1], hist[s][2], hist[s][3]);
        worst = mx > worst ? mx : worst;
        best = mn < best ? mn : best;
    }
    printf("tpb %d blocks/SM %d: max warps on one sub-partition %d, min %d\n", tpb, bps, worst, best);
    for (int w = 0; w < 8 && w < nw; w++) printf("warp %d (block %d): sm %d slot %d\n", w, w / (tpb / 32), h[2 * w], h[2 * w + 1]);
    return 0;
}
