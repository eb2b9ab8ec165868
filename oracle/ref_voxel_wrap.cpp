// C-ABI shim around the REFERENCE's own CPU voxelization, compiled from the
// reference sources where they lie (/root/reference/mmdet3d/ops/voxel/src/
// voxelization_cpu.cpp) into oracle/_ref/libref_voxel.so by oracle/build_ref.py.
// TEST INFRASTRUCTURE ONLY: used to pin oracle/voxel_ops.py and, optionally, as
// the "reference" CPU baseline for the voxelization rows.  No reference source
// is copied into this repository; this file only declares the two entry points
// (signatures from mmdet3d/ops/voxel/src/voxelization.h:14-24,46-56).
#include <torch/torch.h>
#include <vector>

namespace voxelization {
int hard_voxelize_cpu(const at::Tensor& points, at::Tensor& voxels, at::Tensor& coors,
                      at::Tensor& num_points_per_voxel, const std::vector<float> voxel_size,
                      const std::vector<float> coors_range, const int max_points,
                      const int max_voxels, const int NDim);
void dynamic_voxelize_cpu(const at::Tensor& points, at::Tensor& coors,
                          const std::vector<float> voxel_size,
                          const std::vector<float> coors_range, const int NDim);
}  // namespace voxelization

extern "C" {

void ref_dynamic_voxelize_cpu(const float* points, int n, int nfeat, const float* voxel_size,
                              const float* coors_range, int* coors_out) {
  auto p = torch::from_blob(const_cast<float*>(points), {n, nfeat}, torch::kFloat32);
  auto c = torch::from_blob(coors_out, {n, 3}, torch::kInt32);
  voxelization::dynamic_voxelize_cpu(p, c, {voxel_size[0], voxel_size[1], voxel_size[2]},
                                     {coors_range[0], coors_range[1], coors_range[2], coors_range[3],
                                      coors_range[4], coors_range[5]}, 3);
}

int ref_hard_voxelize_cpu(const float* points, int n, int nfeat, const float* voxel_size,
                          const float* coors_range, int max_points, int max_voxels, float* voxels,
                          int* coors, int* num_points_per_voxel) {
  auto p = torch::from_blob(const_cast<float*>(points), {n, nfeat}, torch::kFloat32);
  auto v = torch::from_blob(voxels, {max_voxels, max_points, nfeat}, torch::kFloat32);
  auto c = torch::from_blob(coors, {max_voxels, 3}, torch::kInt32);
  auto k = torch::from_blob(num_points_per_voxel, {max_voxels}, torch::kInt32);
  return voxelization::hard_voxelize_cpu(p, v, c, k, {voxel_size[0], voxel_size[1], voxel_size[2]},
                                         {coors_range[0], coors_range[1], coors_range[2], coors_range[3],
                                          coors_range[4], coors_range[5]}, max_points, max_voxels, 3);
}
}
