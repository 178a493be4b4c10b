// Latency floor of a host-visible step on this box (not the product): how long from a launch until a flag the kernel
// wrote to pinned host memory is seen by the host, for an empty kernel, one with a 4 KB parameter block, one that first
// reads a word of pinned host memory, and the round trip of a doorbell to a resident (persistent) kernel.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_latency tools/probe_latency.cu && /tmp/probe_latency
#include <cstdio>
#include <cstdint>
#include <ctime>
#include <cstring>
#include <cuda_runtime.h>
struct Big { int v[1000]; };
__global__ void k_flag(volatile uint32_t* flag, uint32_t t) { if (threadIdx.x == 0 && blockIdx.x == 0) { *flag = t; } }
__global__ void k_big(Big b, volatile uint32_t* flag, uint32_t t) { if (threadIdx.x == 0 && blockIdx.x == 0) { *flag = t + (b.v[999] & 0); } }
__global__ void k_read(const volatile uint32_t* in, volatile uint32_t* flag, uint32_t t) { if (threadIdx.x == 0 && blockIdx.x == 0) { *flag = t + (*in & 0); } }
__global__ void k_spin(volatile uint32_t* flag, uint32_t t, long long cycles) {
    const long long t0 = clock64(); while (clock64() - t0 < cycles) { }
    if (threadIdx.x == 0 && blockIdx.x == 0) { __threadfence_system(); *flag = t; }
}
// 4,096 warps, each ~7 us of work, then one 96-word record per warp of which 36 words are used: written to `rec` (host or
// device memory) as three unaligned pieces (how the emit rounds of obs_world_compact land) or as coalesced 128-byte rows
__global__ void k_records_t(uint32_t* rec, uint32_t t, long long cycles) {  // every word carries the step's number
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64(); while (clock64() - t0 < cycles) { }
    uint32_t* r = rec + (size_t)w * 96;
    r[lane] = t;
    if (lane < 5) r[32 + lane] = t;
}
__global__ void k_records(uint32_t* rec, int pieces, long long cycles) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const long long t0 = clock64(); while (clock64() - t0 < cycles) { }
    uint32_t* r = rec + (size_t)w * 96;
    if (pieces) {
        if (lane < 4) r[lane] = w;
        if (lane < 15) r[4 + lane] = lane;
        if (lane < 5) r[19 + lane] = lane;
        if (lane < 13) r[24 + lane] = lane;
    } else {
        r[lane] = lane;
        if (lane < 5) r[32 + lane] = lane;
    }
}
__global__ void k_persist(volatile uint32_t* bell, volatile uint32_t* ack, uint32_t last) {
    uint32_t seen = 0;
    while (seen != last) { const uint32_t b = *bell; if (b != seen) { seen = b; __threadfence_system(); *ack = b; } }
}
static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }
static void idle(double us) { const double t = now(); while (now() - t < us) { } }
int main() {
    volatile uint32_t *flag, *in;
    cudaHostAlloc((void**)&flag, 4096, cudaHostAllocMapped); cudaHostAlloc((void**)&in, 4096, cudaHostAllocMapped);
    *flag = 0; *in = 0;
    Big b; for (int& v : b.v) v = 1;
    const int R = 300;
    for (int mode = 0; mode < 5; ++mode) {
        double sum = 0, sum_launch = 0;
        for (int i = 1; i <= R + 20; ++i) {
            idle(60);  // the GPU is idle between steps, as in a host loop
            const double t0 = now();
            const uint32_t t = mode * 10000 + i;
            if (mode == 0) k_flag<<<1, 32>>>(flag, t);
            else if (mode == 1) k_big<<<1, 32>>>(b, flag, t);
            else if (mode == 2) k_read<<<1, 32>>>(in, flag, t);
            else if (mode == 3) k_flag<<<2048, 64>>>(flag, t);
            else k_spin<<<2048, 64>>>(flag, t, 19000);  // ~10 us of work
            const double t1 = now();
            while (*flag != t) { }
            const double t2 = now();
            if (i > 20) { sum += t2 - t0; sum_launch += t1 - t0; }
        }
        const char* names[] = {"empty kernel", "4 KB of parameters", "reads pinned host memory first", "2048 CTAs", "2048 CTAs x 10 us of work"};
        printf("%-34s launch call %.1f us, flag seen after %.1f us\n", names[mode], sum_launch / R, sum / R);
        cudaDeviceSynchronize();
    }
    {   // what does a 16 KB host-to-device copy in front of the kernel cost?  (zs_step_host sends the actions that way)
        uint32_t *hsrc, *ddst;
        cudaHostAlloc((void**)&hsrc, 16384, 0); cudaMalloc((void**)&ddst, 16384);
        struct Acts { unsigned char a[4096]; } acts; memset(&acts, 1, sizeof(acts));
        for (int mode = 0; mode < 2; ++mode) {
            double sum = 0, sum_launch = 0;
            for (int i = 1; i <= R + 20; ++i) {
                idle(60);
                const double t0 = now();
                const uint32_t t = 70000 + mode * 1000 + i;
                if (mode == 1) cudaMemcpyAsync(ddst, hsrc, 16384, cudaMemcpyHostToDevice, 0);
                k_spin<<<2048, 64>>>(flag, t, 19000);
                const double t1 = now();
                while (*flag != t) { }
                if (i > 20) { sum += now() - t0; sum_launch += t1 - t0; }
            }
            printf("%-34s launch calls %.1f us, flag seen after %.1f us\n", mode ? "16 KB H2D copy + 10 us kernel" : "10 us kernel alone", sum_launch / R, sum / R);
            cudaDeviceSynchronize();
        }
    }
    {   // records: host memory in pieces / coalesced, device memory + a copy of the whole buffer
        uint32_t *hrec, *drec, *hcopy;
        const size_t bytes = 4096 * 96 * 4;
        cudaHostAlloc((void**)&hrec, bytes, cudaHostAllocMapped); cudaMalloc((void**)&drec, bytes); cudaHostAlloc((void**)&hcopy, bytes, 0);
        const char* names[] = {"records -> host, 4 pieces", "records -> host, coalesced", "records -> device + 1.5 MB copy", "records -> device + 0.6 MB copy", "records -> device, no copy"};
        for (int mode = 0; mode < 5; ++mode) {
            double sum = 0;
            for (int i = 1; i <= R + 20; ++i) {
                idle(60);
                const double t0 = now();
                const uint32_t t = 50000 + mode * 1000 + i;
                k_records<<<1024, 128>>>(mode <= 1 ? hrec : drec, mode == 0 ? 1 : 0, 13000);
                if (mode == 2) cudaMemcpyAsync(hcopy, drec, bytes, cudaMemcpyDeviceToHost, 0);
                if (mode == 3) cudaMemcpyAsync(hcopy, drec, 4096 * 36 * 4, cudaMemcpyDeviceToHost, 0);
                k_flag<<<1, 32>>>(flag, t);
                while (*flag != t) { }
                if (i > 20) sum += now() - t0;
            }
            printf("%-34s flag seen after %.1f us\n", names[mode], sum / R);
            cudaDeviceSynchronize();
        }
    }
    {   // when do un-fenced record words show up on the host?  (sample warps polled in turn)
        uint32_t* hrec;
        cudaHostAlloc((void**)&hrec, 4096 * 96 * 4, cudaHostAllocMapped);
        memset(hrec, 0, 4096 * 96 * 4);
        const int samples[8] = {0, 1, 512, 1024, 2048, 3072, 4000, 4095};
        double first[8] = {0}, lastw = 0, all = 0;
        for (int i = 1; i <= R + 20; ++i) {
            idle(60);
            const double t0 = now();
            const uint32_t t = 90000 + i;
            k_records_t<<<1024, 128>>>(hrec, t, 13000);
            k_flag<<<1, 32>>>(flag, t);
            double seen[8] = {0};
            int left = 8;
            volatile uint32_t* v = hrec;
            while (left) for (int k = 0; k < 8; ++k) if (!seen[k] && v[(size_t)samples[k] * 96 + 36] == t) { seen[k] = now() - t0; --left; }
            double mx = 0; for (int k = 0; k < 8; ++k) mx = seen[k] > mx ? seen[k] : mx;
            while (*flag != t) { }
            const double tf = now() - t0;
            if (i > 20) { for (int k = 0; k < 8; ++k) first[k] += seen[k]; lastw += mx; all += tf; }
        }
        printf("un-fenced record words seen after (us), warps 0/1/512/1024/2048/3072/4000/4095:");
        for (int k = 0; k < 8; ++k) printf(" %.1f", first[k] / R);
        printf("; last of them %.1f; flag kernel %.1f\n", lastw / R, all / R);
    }
    // stream sync instead of a flag
    { double sum = 0; for (int i = 1; i <= R; ++i) { idle(60); const double t0 = now(); k_flag<<<1, 32>>>(flag, 7); cudaStreamSynchronize(0); sum += now() - t0; }
      printf("%-34s %.1f us\n", "empty kernel + cudaStreamSynchronize", sum / R); }
    // doorbell round trip of a resident kernel
    volatile uint32_t *bell, *ack;
    cudaHostAlloc((void**)&bell, 4096, cudaHostAllocMapped); cudaHostAlloc((void**)&ack, 4096, cudaHostAllocMapped);
    *bell = 0; *ack = 0;
    k_persist<<<1, 1>>>(bell, ack, R + 1);
    double sum = 0;
    for (uint32_t i = 1; i <= (uint32_t)R + 1; ++i) {
        idle(60);
        const double t0 = now();
        *bell = i; __sync_synchronize();
        double tw = now();
        while (*ack != i) { if (now() - tw > 2e6) { printf("doorbell: no answer\n"); return 1; } }
        if (i > 1) sum += now() - t0;
    }
    cudaDeviceSynchronize();
    printf("%-34s %.1f us\n", "doorbell round trip (resident kernel)", sum / R);
    return 0;
}
