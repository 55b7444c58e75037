// vio_host.h — small host-side RAII helpers shared by the C-ABI translation unit.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

template <typename T>
struct DBuf {  // device buffer
    T *p = nullptr;
    size_t n = 0;    // logical element count
    size_t cap = 0;  // allocated element count: a handle that is re-used for many graphs (vio_solve_batched) only
                     // re-allocates when a buffer has to grow
    cudaError_t alloc(size_t count) {
        n = count;
        if (count <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = count + count / 4 + 16;
        cudaError_t e = cudaMalloc((void **)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want; else n = 0;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        n = 0;
        cap = 0;
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
