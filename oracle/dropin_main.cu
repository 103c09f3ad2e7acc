// dropin_main.cu -- link-level drop-in check.  TEST INFRASTRUCTURE ONLY (our code, not reference code).
//
// oracle/Makefile (target `dropin`) compiles the REFERENCE's own src/world/octree.cpp against the REFERENCE's own
// headers and glm, unmodified, and links that object -- together with this driver, which is compiled against the
// reference's headers as well -- against libosl_host.so / libosl_b200.so instead of the reference's svo.cu /
// cone_tracing_kernels.cu / common_types.cu.  If the seam's linkage (extern "C", svo.h:14-18,
// cone_tracing_kernels.h:16) or calling convention (glm 0.9.5 types by value) differed, this would not link or the
// arguments would arrive garbled; tests/test_host_shim.py compares what it writes with the CPU oracle.
//
//   ref_octree_dropin in.bin out.bin
//   in.bin : int n, w, h; float center[3], size, resolution, fov, view[16]; n * vec3 points; n * Color256
//   out.bin: int n_nodes; 2*n_nodes uint pool; w*h uchar4 image; int n_voxels; n_voxels * vec4 centres, colours
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define private public  // the test reads OctreeNode::gpu_size_ (the reference exposes no node count)
#include <octree_slam/world/octree.h>
#undef private
#include <octree_slam/rendering/cone_tracing_kernels.h>

using namespace octree_slam;

struct Header {
  int n, w, h;
  float center[3], size, resolution, fov, view[16];
};

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  Header hd;
  if (fread(&hd, sizeof(hd), 1, f) != 1) return 4;
  std::vector<glm::vec3> pts(hd.n);
  std::vector<Color256> cols(hd.n);
  if (fread(pts.data(), sizeof(glm::vec3), hd.n, f) != (size_t)hd.n) return 4;
  if (fread(cols.data(), sizeof(Color256), hd.n, f) != (size_t)hd.n) return 4;
  fclose(f);
  glm::vec3* d_pts;
  Color256* d_cols;
  cudaMalloc((void**)&d_pts, sizeof(glm::vec3) * hd.n);
  cudaMalloc((void**)&d_cols, sizeof(Color256) * hd.n);
  cudaMemcpy(d_pts, pts.data(), sizeof(glm::vec3) * hd.n, cudaMemcpyHostToDevice);
  cudaMemcpy(d_cols, cols.data(), sizeof(Color256) * hd.n, cudaMemcpyHostToDevice);

  const glm::vec3 center(hd.center[0], hd.center[1], hd.center[2]);
  world::Octree tree(hd.resolution, center, hd.size);  // the reference's class, from the reference's octree.cpp
  BoundingBox bbox = tree.boundingBox();
  tree.addCloud(glm::vec3(0.0f), d_pts, d_cols, hd.n, bbox);  // -> svo::svoFromPointCloud (octree.cpp:290)
  tree.addCloud(glm::vec3(0.0f), d_pts, d_cols, hd.n, bbox);
  SVO svo = tree.extractSVO(bbox);                             // octree.cpp:339-360
  const int n_nodes = tree.root_->gpu_size_;

  uchar4* d_img;
  cudaMalloc((void**)&d_img, sizeof(uchar4) * hd.w * hd.h);
  glm::mat4 view;
  for (int c = 0; c < 4; c++)
    for (int r = 0; r < 4; r++) view[c][r] = hd.view[4 * c + r];
  rendering::coneTraceSVO(d_img, glm::vec2((float)hd.w, (float)hd.h), hd.fov, view, svo);  // cuda_renderer.cpp:163

  VoxelGrid grid;
  grid.bbox = bbox;
  grid.scale = hd.resolution;
  tree.extractVoxelGrid(grid);                                 // -> svo::extractVoxelGridFromSVO (octree.cpp:336)

  std::vector<unsigned int> pool(2 * (size_t)n_nodes);
  std::vector<uchar4> img((size_t)hd.w * hd.h);
  std::vector<glm::vec4> cen(grid.size), col(grid.size);
  cudaMemcpy(pool.data(), svo.data, sizeof(unsigned int) * pool.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(img.data(), d_img, sizeof(uchar4) * img.size(), cudaMemcpyDeviceToHost);
  if (grid.size > 0) {
    cudaMemcpy(cen.data(), grid.centers, sizeof(glm::vec4) * grid.size, cudaMemcpyDeviceToHost);
    cudaMemcpy(col.data(), grid.colors, sizeof(glm::vec4) * grid.size, cudaMemcpyDeviceToHost);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return 5;
  f = fopen(argv[2], "wb");
  if (!f) return 3;
  fwrite(&n_nodes, sizeof(int), 1, f);
  fwrite(pool.data(), sizeof(unsigned int), pool.size(), f);
  fwrite(img.data(), sizeof(uchar4), img.size(), f);
  fwrite(&grid.size, sizeof(int), 1, f);
  fwrite(cen.data(), sizeof(glm::vec4), cen.size(), f);
  fwrite(col.data(), sizeof(glm::vec4), col.size(), f);
  fclose(f);
  cudaFree(d_pts); cudaFree(d_cols); cudaFree(d_img);
  printf("dropin ok: %d nodes, %d voxels\n", n_nodes, grid.size);
  return 0;
}
