// ipc_probe.cu -- can two PROCESSES share device memory through CUDA IPC on this box (same GPU, and GPU 0 <-> GPU 1),
// with one side's kernel writing records + a flag into the other's buffer while the other side's kernel spins on the flag?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o ipc_probe ipc_probe.cu && ./ipc_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/wait.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("[%d] %s: %s\n", getpid(), #x, cudaGetErrorString(e_)); exit(2); } } while (0)

__global__ void k_push(double *peer_buf, volatile int *peer_flag, int n, int epoch)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) peer_buf[i] = (double)(i + epoch);
    __threadfence_system();
    __shared__ int last;
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd((int *)(peer_flag + 1), 1) == (int)gridDim.x - 1;   // counter lives next to the flag
    __syncthreads();
    if (last && threadIdx.x == 0) { __threadfence_system(); peer_flag[1] = 0; __threadfence_system(); peer_flag[0] = epoch; }
}
__global__ void k_wait(volatile int *flag, int epoch, int *timed_out)
{
    long long t0 = clock64();
    while (*flag < epoch) { if (clock64() - t0 > 20000000000LL) { *timed_out = 1; break; } __nanosleep(200); }
}
__global__ void k_check(const double *buf, int n, int epoch, int *bad)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) if (buf[i] != (double)(i + epoch)) atomicAdd(bad, 1);
}

static int role(bool child, int devA, int devB, int rd, int wr)
{
    const int n = 1 << 20;
    const int dev = child ? devB : devA;
    CK(cudaSetDevice(dev));
    double *buf; int *flag, *res;
    CK(cudaMalloc(&buf, n * sizeof(double) + 256));
    flag = (int *)(buf + n);
    CK(cudaMemset(buf, 0, n * sizeof(double) + 256));
    CK(cudaMalloc(&res, 8)); CK(cudaMemset(res, 0, 8));
    cudaIpcMemHandle_t mine, theirs;
    CK(cudaIpcGetMemHandle(&mine, buf));
    if (write(wr, &mine, sizeof mine) != sizeof mine || read(rd, &theirs, sizeof theirs) != sizeof theirs) return 4;
    double *peer;
    CK(cudaIpcOpenMemHandle((void **)&peer, theirs, cudaIpcMemLazyEnablePeerAccess));
    int *peer_flag = (int *)(peer + n);
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float ms = 0;
    for (int epoch = 1; epoch <= 20; epoch++) {
        if (epoch == 11) CK(cudaEventRecord(e0, st));
        k_push<<<148, 256, 0, st>>>(peer, peer_flag, n, epoch);
        k_wait<<<1, 1, 0, st>>>(flag, epoch, res + 1);
        k_check<<<148, 256, 0, st>>>(buf, n, epoch, res);
        CK(cudaStreamSynchronize(st));          // host handshake: the partner must not overwrite my buffer before I checked it
        char c = 1;
        if (write(wr, &c, 1) != 1 || read(rd, &c, 1) != 1) return 5;
    }
    CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    int h[2]; CK(cudaMemcpy(h, res, 8, cudaMemcpyDeviceToHost));
    printf("[%s dev %d -> peer dev %d] 20 epochs of 8 MiB push+flag: bad=%d timed_out=%d, last 10 epochs %.3f ms each (incl. host handshake)\n",
           child ? "B" : "A", dev, child ? devA : devB, h[0], h[1], ms / 10);
    CK(cudaIpcCloseMemHandle(peer));
    return (h[0] || h[1]) ? 1 : 0;
}

// the launcher never touches CUDA: both roles are forked children with their own contexts
static int run_pair(int devA, int devB)
{
    int a2b[2], b2a[2];
    if (pipe(a2b) || pipe(b2a)) return 3;
    pid_t pa = fork();
    if (pa == 0) exit(role(false, devA, devB, b2a[0], a2b[1]));
    pid_t pb = fork();
    if (pb == 0) exit(role(true, devA, devB, a2b[0], b2a[1]));
    int sa = 0, sb = 0;
    waitpid(pa, &sa, 0); waitpid(pb, &sb, 0);
    return (WIFEXITED(sa) && WEXITSTATUS(sa) == 0 && WIFEXITED(sb) && WEXITSTATUS(sb) == 0) ? 0 : 1;
}

int main(int argc, char **argv)
{
    const int ndev = argc > 1 ? atoi(argv[1]) : 1;
    int rc = run_pair(0, 0);
    printf("same-GPU IPC: %s\n", rc ? "FAILED" : "ok");
    if (ndev > 1) { int r2 = run_pair(0, 1); printf("GPU0<->GPU1 IPC: %s\n", r2 ? "FAILED" : "ok"); rc |= r2; }
    return rc;
}
