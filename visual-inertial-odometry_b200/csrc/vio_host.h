// vio_host.h — small host-side RAII helpers shared by the C-ABI translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

template <typename T>
struct DBuf {  // device buffer
    T *p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void **)&p, count * sizeof(T));
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
    }
    DBuf() = default;
    DBuf(const DBuf &) = delete;
    DBuf &operator=(const DBuf &) = delete;
    ~DBuf() { release(); }
};

template <typename T>
inline cudaError_t upload(DBuf<T> &d, const T *h, size_t n, cudaStream_t s) {
    cudaError_t e = d.alloc(n);
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpyAsync(d.p, h, n * sizeof(T), cudaMemcpyHostToDevice, s);
}
