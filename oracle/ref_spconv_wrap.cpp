// C-ABI shim around the REFERENCE's vendored spconv-1.x CPU rulebook generator, compiled from the header where it
// lies (/root/reference/mmdet3d/ops/bevfusion-ops/spconv/include/spconv/geometry.h:144-297, getIndicePairsConv /
// getIndicePairsSubM) into oracle/_ref/libref_spconv.so by oracle/build_ref.py.
// TEST INFRASTRUCTURE ONLY: pins oracle/spconv_ops.py::get_indice_pairs (tests/test_oracle_spconv.py).  No
// reference source is copied into this repository; the buffers are set up the way getIndicePair does it
// (spconv_ops.h:50-137: indicePairs filled with -1, grid with -1, SubM uses stride 1 / padding k/2).
#include <spconv/geometry.h>

#include <vector>

extern "C" {

// indices_in [n,4] (b,z,y,x) int32; out_shape[3]; pairs [kvol,2,n] (pre-filled -1 by the caller); num [kvol] zeroed;
// indices_out [n*kvol,4] (strided conv only).  Returns the number of active output sites.
int ref_spconv_indice_pairs(const int* indices_in, int n, int batch, const int* ksize, const int* stride,
                            const int* padding, const int* dilation, const int* out_shape, int subm,
                            int* indices_out, int* pairs, int* num) {
  int kvol = ksize[0] * ksize[1] * ksize[2];
  long vol = 1L * out_shape[0] * out_shape[1] * out_shape[2];
  std::vector<int> grid(static_cast<size_t>(batch) * vol, -1);
  tv::TensorView<const int> in(indices_in, {n, 4});
  tv::TensorView<int> g(grid.data(), {static_cast<int>(batch * vol)});
  tv::TensorView<int> p(pairs, {kvol, 2, n});
  tv::TensorView<int> c(num, {kvol});
  if (subm) {
    int st[3] = {1, 1, 1};
    int pd[3] = {ksize[0] / 2, ksize[1] / 2, ksize[2] / 2};
    return spconv::getIndicePairsSubM<int, int, 3>(in, g, p, c, ksize, st, pd, dilation, out_shape);
  }
  tv::TensorView<int> out(indices_out, {n * kvol, 4});
  return spconv::getIndicePairsConv<int, int, 3>(in, out, g, p, c, ksize, stride, padding, dilation, out_shape);
}
}
