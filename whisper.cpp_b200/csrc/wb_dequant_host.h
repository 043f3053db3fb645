// wb_dequant_host.h -- ggml block formats this engine has no device kernels for (Q4_1, Q5_1, Q2_K, Q3_K, Q6_K, and BF16 values) are expanded on the
// HOST at load time and live in HBM as F16 matrices (the validated F16 path runs them).  The reference never materialises these values
// -- its CPU path multiplies the blocks with Q8_1 / Q8_K activation blocks -- so this trades HBM (2 bytes per weight) for coverage of
// the checkpoints that are distributed in these formats (e.g. the q5_1 tiny/base/small files).
// Block layouts: ggml/src/ggml-common.h:176-330; semantics: dequantize_row_* in ggml/src/ggml-quants.c.
#pragma once
#include <cstddef>
#include <cstdint>

namespace wb {

enum HostDqType : int { HT_Q4_1 = 3, HT_Q5_1 = 7, HT_Q2_K = 10, HT_Q3_K = 11, HT_Q6_K = 14, HT_BF16 = 30, HT_Q4_K = 12, HT_Q5_K = 13 };   // ggml_type ids

bool   host_dq_supported(int ggml_type);
int    host_dq_block_values(int ggml_type);        // 32 or 256 (1 for BF16)
size_t host_dq_block_bytes(int ggml_type);
// n values (a multiple of the block size) from `src` blocks to f32
void   host_dequantize(int ggml_type, const void * src, float * dst, int64_t n);

} // namespace wb
