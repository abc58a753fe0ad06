// host_pipe.cuh — double-buffered staging behind the host-pointer entry points (wc_*_tick_host_async / _sync).
// Three streams per handle (upload, kernels, download) and three staging slots: the upload of tick t+1 and the
// download of tick t-1 overlap the kernels of tick t.  Up to three ticks may be in flight; results of a tick are
// valid after the wc_*_sync_one / wc_*_sync call that retires it.
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace wcb {

struct HostPipe {
    static constexpr int kSlots = 3;
    cudaStream_t up = nullptr, run = nullptr, run2 = nullptr, down = nullptr; // run2: a second kernel stream (fused MPC + IK tick)
    cudaEvent_t fork = nullptr, join = nullptr;
    size_t in_cap = 0, out_cap = 0;
    void* dev_in[kSlots] = {nullptr, nullptr};
    void* dev_out[kSlots] = {nullptr, nullptr};
    cudaEvent_t up_done[kSlots] = {nullptr, nullptr}, run_done[kSlots] = {nullptr, nullptr}, down_done[kSlots] = {nullptr, nullptr};
    bool used[kSlots] = {false, false};
    int64_t submitted = 0, retired = 0;
    bool ready = false;

    cudaError_t init(size_t in_bytes, size_t out_bytes)
    {
        if (ready && in_bytes <= in_cap && out_bytes <= out_cap) return cudaSuccess;
        if (ready) { // a larger entry point on the same handle (e.g. the fused MPC + IK tick): drain and re-create
            const cudaError_t es = sync_all();
            if (es != cudaSuccess) return es;
            destroy();
        }
        const cudaError_t e = init_once(in_bytes, out_bytes);
        if (e != cudaSuccess) destroy(); // a partial set of streams / buffers must not leak into the next attempt
        return e;
    }
    cudaError_t init_once(size_t in_bytes, size_t out_bytes)
    {
        cudaError_t e;
        if ((e = cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&run, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&run2, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaStreamCreateWithFlags(&down, cudaStreamNonBlocking)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&join, cudaEventDisableTiming)) != cudaSuccess) return e;
        for (int k = 0; k < kSlots; k++) {
            if ((e = cudaMalloc(&dev_in[k], in_bytes)) != cudaSuccess) return e;
            if ((e = cudaMalloc(&dev_out[k], out_bytes)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&up_done[k], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&run_done[k], cudaEventDisableTiming)) != cudaSuccess) return e;
            if ((e = cudaEventCreateWithFlags(&down_done[k], cudaEventDisableTiming)) != cudaSuccess) return e;
        }
        in_cap = in_bytes; out_cap = out_bytes;
        ready = true;
        return cudaSuccess;
    }
    void destroy()
    {
        for (int k = 0; k < kSlots; k++) {
            if (dev_in[k]) cudaFree(dev_in[k]);
            if (dev_out[k]) cudaFree(dev_out[k]);
            if (up_done[k]) cudaEventDestroy(up_done[k]);
            if (run_done[k]) cudaEventDestroy(run_done[k]);
            if (down_done[k]) cudaEventDestroy(down_done[k]);
            dev_in[k] = dev_out[k] = nullptr; up_done[k] = run_done[k] = down_done[k] = nullptr; used[k] = false;
        }
        if (up) cudaStreamDestroy(up);
        if (run) cudaStreamDestroy(run);
        if (run2) cudaStreamDestroy(run2);
        if (down) cudaStreamDestroy(down);
        if (fork) cudaEventDestroy(fork);
        if (join) cudaEventDestroy(join);
        up = run = run2 = down = nullptr; fork = join = nullptr;
        submitted = retired = 0; in_cap = out_cap = 0;
        ready = false;
    }
    // retire the oldest tick in flight (its outputs are then in the caller's buffers)
    cudaError_t sync_one()
    {
        if (retired >= submitted) return cudaSuccess;
        const cudaError_t e = cudaEventSynchronize(down_done[retired % kSlots]);
        retired++;
        return e;
    }
    cudaError_t sync_all()
    {
        cudaError_t e = cudaSuccess;
        while (retired < submitted && e == cudaSuccess) e = sync_one();
        return e;
    }
    // slot of the next tick; waits on the host only if all slots are still in flight
    int begin(cudaError_t* err)
    {
        *err = cudaSuccess;
        if (submitted - retired >= kSlots) *err = sync_one();
        const int k = (int)(submitted % kSlots);
        if (used[k] && *err == cudaSuccess) *err = cudaStreamWaitEvent(up, run_done[k], 0); // kernels of tick t-2 no longer read dev_in[k]
        return k;
    }
    // call after the uploads were enqueued on `up`: the kernels wait for them and for the slot's previous download
    cudaError_t before_run(int k)
    {
        cudaError_t e = cudaEventRecord(up_done[k], up);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(run, up_done[k], 0);
        if (e == cudaSuccess && used[k]) e = cudaStreamWaitEvent(run, down_done[k], 0);
        return e;
    }
    // call after the kernels were enqueued on `run`: the downloads wait for them
    cudaError_t before_down(int k)
    {
        cudaError_t e = cudaEventRecord(run_done[k], run);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(down, run_done[k], 0);
        return e;
    }
    cudaError_t end(int k)
    {
        const cudaError_t e = cudaEventRecord(down_done[k], down);
        used[k] = true;
        submitted++;
        return e;
    }
};

} // namespace wcb
