// Host-side construction of TMA tensor maps (cuTensorMapEncodeTiled resolved at run time through the CUDA
// runtime's driver-entry-point API, so the library has no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace hv {

// fp16 row-major matrix [rows][ld] of which `cols` columns are valid; box = box_rows x 64 columns, 128B swizzle.
// Out-of-bounds box elements read as zero.
bool make_map_2d(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows);

// Same with an explicit box width (<= 64 columns).
bool make_map_2d_box(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);

// channels-last activation (N, H, W, C) fp16 -> 4-D map (C, W, H, N), box (64, bw, bh, bn).
bool make_map_nhwc(CUtensorMap* m, const void* ptr, int64_t N, int64_t H, int64_t W, int64_t C, int bn, int bh, int bw);

// channels-last activation viewed as (N, H/2, 2, W/2, 2C) -> 5-D map, box (64, bw, 1, bh, bn): the operand view of a
// stride-2 3x3 convolution (H and W must be even).
bool make_map_nhwc_s2(CUtensorMap* m, const void* ptr, int64_t N, int64_t H, int64_t W, int64_t C, int bn, int bh, int bw);

// fp16 [batch][rows][ld] (cols valid) -> 3-D map (cols, rows, batch), box (64, box_rows, 1); rows beyond `rows` read as zero
// even when the next batch item follows in memory.
bool make_map_3d(CUtensorMap* m, const void* ptr, int64_t batch, int64_t rows, int64_t cols, int64_t ld, int box_rows);

// Epilogue I/O box of the GEMM (TMA store of the output tile / TMA load of the residual tile): fp16 [rows][ld] with `cols`
// valid columns, box = box_rows x box_cols, shared-memory rows of box_cols * 2 bytes (128 -> 128B swizzle, 64 -> 64B swizzle,
// anything else unswizzled).  Out-of-range rows / columns are clipped on store and read as zero on load.
// The same for the batched-B GEMM (V^T producer): out[row][b * batch_stride + col], col < cols: 3-D map (cols, rows, batch), box
// (box_cols, box_rows, 1).  Columns beyond `cols` are clipped per batch item, so a frame's ragged last tile never spills into the next frame.
bool make_map_3d_io(CUtensorMap* m, const void* ptr, int64_t cols, int64_t rows, int64_t batch, int64_t row_stride, int64_t batch_stride, int box_rows,
                    int box_cols);
bool make_map_2d_io(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols);

// Token matrix [frames][pixels][ld] (cols valid) viewed along the frame axis: 3-D map (cols, pixels, frames), unswizzled box
// (box_cols, 1, box_frames) = the box_frames rows of ONE pixel, box_cols channels wide (temporal attention operands).
bool make_map_frames(CUtensorMap* m, const void* ptr, int64_t frames, int64_t pixels, int64_t cols, int64_t ld, int box_cols, int box_frames);

const char* tma_last_error();

}  // namespace hv
