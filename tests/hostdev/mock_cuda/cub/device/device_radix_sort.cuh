// TEST INFRASTRUCTURE (see ../../mock_runtime.h): the two pieces of cub the host glue uses, on the host.
// DeviceRadixSort::SortPairs with a DoubleBuffer: stable, ascending, on key bits [begin_bit, end_bit).
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

namespace cub {
template <class T>
struct DoubleBuffer {
    T* d_buffers[2];
    int selector = 0;
    DoubleBuffer(T* current, T* alternate) { d_buffers[0] = current; d_buffers[1] = alternate; }
    T* Current() const { return d_buffers[selector]; }
    T* Alternate() const { return d_buffers[selector ^ 1]; }
};
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void* temp, size_t& temp_bytes, DoubleBuffer<K>& keys, DoubleBuffer<V>& vals, int n, int begin_bit, int end_bit,
                                 cudaStream_t) {
        if (!temp) { temp_bytes = 64; return cudaSuccess; }
        const K mask = end_bit >= (int)(8 * sizeof(K)) ? ~(K)0 : (((K)1 << end_bit) - 1);
        std::vector<int> perm(n);
        std::iota(perm.begin(), perm.end(), 0);
        const K* k = keys.Current();
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return ((k[a] & mask) >> begin_bit) < ((k[b] & mask) >> begin_bit); });
        K* ko = keys.Alternate();
        V* vo = vals.Alternate();
        const V* v = vals.Current();
        for (int i = 0; i < n; i++) { ko[i] = k[perm[i]]; vo[i] = v[perm[i]]; }
        keys.selector ^= 1;  // like the library: the result is in the other buffer, the old one is scratch
        vals.selector ^= 1;
        return cudaSuccess;
    }
};
}  // namespace cub
