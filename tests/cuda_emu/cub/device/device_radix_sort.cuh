// stub of <cub/device/device_radix_sort.cuh> for tests/cuda_emu: a stable sort on the selected key bits
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>
namespace cub {
struct DeviceRadixSort {
    template <class K, class V>
    static cudaError_t SortPairs(void* tmp, size_t& need, const K* kin, K* kout, const V* vin, V* vout, int n,
                                 int begin_bit, int end_bit, cudaStream_t = nullptr)
    {
        if (!tmp) { need = 256; return cudaSuccess; }
        const K mask = (end_bit - begin_bit >= (int)(8 * sizeof(K))) ? ~K(0) : ((K(1) << (end_bit - begin_bit)) - 1);
        std::vector<int> idx(n);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return ((kin[a] >> begin_bit) & mask) < ((kin[b] >> begin_bit) & mask); });
        for (int i = 0; i < n; ++i) { kout[i] = kin[idx[i]]; vout[i] = vin[idx[i]]; }
        return cudaSuccess;
    }
};
}  // namespace cub
