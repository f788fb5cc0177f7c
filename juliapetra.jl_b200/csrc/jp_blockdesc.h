// jp_blockdesc.h -- one row block of the SpMV/SpMM kernels (shared by the host runtime and the kernel headers)
#pragma once
#include <cstddef>
#include <cstdint>

namespace jp {

// 32 bytes, read by the kernels as two int4
struct alignas(16) BlockDesc {
  int32_t row0;         // first row
  int32_t nrows_kind;   // nrows | kind << 16
  int32_t p0, p1;       // nonzero range [p0, p1)
  int32_t ucol0;        // SpMM tile path: offset of the block's unique-column list in Csr::ucols (multiple of 4)
  int32_t ucount;       // number of unique columns of the block
  int32_t pad0, pad1;
};
constexpr size_t kBlockDescBytes = sizeof(BlockDesc);

}  // namespace jp
