// cub_ab.cu -- A/B baseline only, NOT part of the product library: cub::DeviceRadixSort::SortPairs
// on (u64 key, u32 value) pairs, as the reference's rasterizer calls it ([upstream]
// rasterizer_impl.cu).  bench.py times it beside sgs_sort_pairs_u64 on the
// same keys (the `ab.sort` entry of the bench line).  Built by __graft_entry__.build() into
// tools/_build/libcub_ab.so with the same nvcc flags as the product.
#include <cub/cub.cuh>
#include <cuda_runtime.h>

extern "C" int cub_ab_sort_pairs_u64(const unsigned long long* keys_in, unsigned long long* keys_out,
                                     const unsigned int* vals_in, unsigned int* vals_out, long long n,
                                     int end_bit, void* temp, size_t* temp_bytes, void* stream) {
    // temp == null: *temp_bytes receives the scratch size (CUB's two-phase convention)
    return (int)cub::DeviceRadixSort::SortPairs(temp, *temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0,
                                                end_bit, (cudaStream_t)stream);
}
