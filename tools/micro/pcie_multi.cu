// micro-benchmark: the host -> device (and device -> host) copy ceiling of ONE box when several GPUs copy at once.
// The e2e number of bench.py is bound by exactly this leg (pinned host cu8 -> HBM, 2 bytes per IQ sample), so its
// scaling over 1/2/4/8 GPUs can be no better than what this prints.
//
//   nvcc -O2 -o pcie_multi pcie_multi.cu -lpthread
//   ./pcie_multi [-m MiB per copy] [-i copies] [-k flat|2d|d2h|wc] [-n numa node for the host buffers | -n -2 = GPU's own]
//                [-p pin thread g to core g*stride] SET [SET ...]        SET = comma-separated GPU ordinals, e.g. 0,1,2,3
//
// One host thread per GPU of a set; every thread allocates its own pinned buffer (after the optional CPU / NUMA binding),
// all threads start together, each queues `copies` asynchronous copies and waits; the set's aggregate rate is total bytes
// over the wall time from the common start to the last thread's finish, the per-GPU rates are CUDA-event times.
#include <cuda_runtime.h>
#include <pthread.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>
#include <vector>

static size_t g_mb = 1024;
static int g_iters = 6, g_numa = -1, g_pin = -1;
static const char *g_kind = "flat";

struct job {
    int dev, idx, n;
    pthread_barrier_t *bar;
    double gbs, t_end;
    int numa_of_dev;
};

static double now() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static int numa_node_of(int dev)
{
    char bus[64], path[256];
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) != cudaSuccess) return -1;
    for (char *c = bus; *c; c++) if (*c >= 'A' && *c <= 'Z') *c += 32;
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    int node = -1;
    if (f) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
    return node;
}

static void *worker(void *arg)
{
    job *j = (job *)arg;
    cudaSetDevice(j->dev);
    if (g_pin >= 0) {
        cpu_set_t cs; CPU_ZERO(&cs); CPU_SET((j->idx * (g_pin ? g_pin : 1)) % (int)sysconf(_SC_NPROCESSORS_ONLN), &cs);
        sched_setaffinity(0, sizeof(cs), &cs);
    }
    int node = g_numa == -2 ? j->numa_of_dev : g_numa;
    if (node >= 0) {   /* MPOL_BIND = 2: pages of this thread's next allocations come from `node` */
        unsigned long mask[16]; memset(mask, 0, sizeof(mask)); mask[node / 64] |= 1ul << (node % 64);
        if (syscall(SYS_set_mempolicy, 2, mask, 1024) != 0) perror("set_mempolicy");
    }
    const size_t bytes = g_mb << 20, rows = g_mb, row_bytes = 1 << 20, dpitch = row_bytes + 1280;
    unsigned char *h = nullptr, *d = nullptr;
    const bool wc = !strcmp(g_kind, "wc"), d2h = !strcmp(g_kind, "d2h"), twod = !strcmp(g_kind, "2d");
    if (cudaHostAlloc(&h, bytes, wc ? cudaHostAllocWriteCombined : cudaHostAllocDefault) != cudaSuccess) { fprintf(stderr, "cudaHostAlloc failed\n"); exit(1); }
    if (!wc) memset(h, 1, bytes);
    if (cudaMalloc(&d, twod ? rows * dpitch : bytes) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); exit(1); }
    cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto copy = [&]() {
        if (d2h) cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s);
        else if (twod) cudaMemcpy2DAsync(d + 1024, dpitch, h, row_bytes, row_bytes, rows, cudaMemcpyHostToDevice, s);
        else cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s);
    };
    copy(); cudaStreamSynchronize(s);                 /* warm-up: page tables, first touch */
    pthread_barrier_wait(j->bar);
    cudaEventRecord(e0, s);
    for (int i = 0; i < g_iters; i++) copy();
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    j->t_end = now();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    j->gbs = (double)bytes * g_iters / (ms * 1e-3) / 1e9;
    pthread_barrier_wait(j->bar);
    cudaFreeHost(h); cudaFree(d);
    return nullptr;
}

int main(int argc, char **argv)
{
    int a = 1;
    for (; a < argc && argv[a][0] == '-' && argv[a][1] && !(argv[a][1] >= '0' && argv[a][1] <= '9'); a += 2) {
        if (a + 1 >= argc) return 2;
        if (!strcmp(argv[a], "-m")) g_mb = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "-i")) g_iters = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "-k")) g_kind = argv[a + 1];
        else if (!strcmp(argv[a], "-n")) g_numa = atoi(argv[a + 1]);
        else if (!strcmp(argv[a], "-p")) g_pin = atoi(argv[a + 1]);
        else return 2;
    }
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    for (; a < argc; a++) {
        std::vector<int> devs;
        char *dup = strdup(argv[a]);
        for (char *t = strtok(dup, ","); t; t = strtok(nullptr, ",")) if (atoi(t) < ndev) devs.push_back(atoi(t));
        free(dup);
        if (devs.empty()) continue;
        const int n = (int)devs.size();
        pthread_barrier_t bar; pthread_barrier_init(&bar, nullptr, n + 1);
        std::vector<job> jobs(n); std::vector<pthread_t> th(n);
        for (int i = 0; i < n; i++) { jobs[i] = {devs[i], i, n, &bar, 0, 0, numa_node_of(devs[i])}; pthread_create(&th[i], nullptr, worker, &jobs[i]); }
        pthread_barrier_wait(&bar);
        const double t0 = now();
        pthread_barrier_wait(&bar);
        double t1 = t0;
        for (int i = 0; i < n; i++) if (jobs[i].t_end > t1) t1 = jobs[i].t_end;
        for (int i = 0; i < n; i++) pthread_join(th[i], nullptr);
        printf("{\"kind\": \"%s\", \"numa\": %d, \"pin\": %d, \"gpus\": \"%s\", \"n\": %d, \"aggregate_gbs\": %.1f, \"per_gpu_gbs\": [", g_kind, g_numa, g_pin, argv[a], n,
               (double)(g_mb << 20) * g_iters * n / (t1 - t0) / 1e9);
        for (int i = 0; i < n; i++) printf("%s%.1f", i ? ", " : "", jobs[i].gbs);
        printf("], \"gpu_numa\": [");
        for (int i = 0; i < n; i++) printf("%s%d", i ? ", " : "", jobs[i].numa_of_dev);
        printf("]}\n");
        fflush(stdout);
        pthread_barrier_destroy(&bar);
    }
    return 0;
}
