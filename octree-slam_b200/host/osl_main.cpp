// osl_main.cpp -- headless replay of the reference's frame loop (src/main.cpp:31-62) on top of the drop-in headers.
// The body of the loop is the reference's call sequence verbatim (main.cpp:38-44, 56-58); what differs is what the
// reference cannot do headless: frames come from a file instead of an OpenNI camera (openni_device.cpp:96-150), the
// pose comes with the frame instead of RGBDCamera (whose update is commented out, main.cpp:35), and the renderer
// writes a device buffer instead of a GL PBO.
//
//   osl_main <frames.bin> <out_prefix> [fused|track|slam]
// track: main.cpp:35 uncommented -- the pose of every frame comes from sensor::RGBDCamera (corrected tracker) instead
//        of the file; the estimated poses are written to <out_prefix>.poses (n x 16 floats, column-major).
// slam:  track, and every frame after the first is fused with the pose read ON THE DEVICE (Octree::addDepthFrame
//        with the camera): tracker and integration of a frame are 31 launches without a host round trip.
// frames.bin: int32 w, h, n; float fx, fy; then n x { float pose[16] (column-major), uint16 depth[w*h], uint8 rgb[w*h*3] }
// writes <out_prefix>.pool (int32 n_nodes, float center[3], float half, uint32 pool[2n]) and <out_prefix>.rgba (w*h*4).
#include <cuda_runtime_api.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <octree_slam/common_types.h>
#include <octree_slam/rendering/cuda_renderer.h>
#include <octree_slam/sensor/image_kernels.h>
#include <octree_slam/sensor/rgbd_camera.h>
#include <octree_slam/world/scene.h>

using namespace octree_slam;

static bool read_exact(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

// osl_main mesh <mesh.bin> <out_prefix>: Scene::voxelizeMeshes(true) (scene.cpp:64-85) on a mesh file
// mesh.bin: int32 nv, nt; float vertices[3*nv]; int32 indices[3*nt]   (bbox = min/max of the vertices)
static int mesh_main(const char* mesh_path, const char* out_prefix) {
  FILE* f = fopen(mesh_path, "rb");
  if (!f) { perror(mesh_path); return 2; }
  int hdr[2];
  if (!read_exact(f, hdr, sizeof(hdr))) return 2;
  std::vector<float> vbo((size_t)hdr[0] * 3);
  std::vector<int> ibo((size_t)hdr[1] * 3);
  if (!read_exact(f, vbo.data(), vbo.size() * 4) || !read_exact(f, ibo.data(), ibo.size() * 4)) return 2;
  fclose(f);
  Mesh mesh;
  mesh.vbo = vbo.data(); mesh.vbosize = (int)vbo.size();
  mesh.ibo = ibo.data(); mesh.ibosize = (int)ibo.size();
  mesh.bbox.bbox0 = glm::vec3(vbo[0], vbo[1], vbo[2]);
  mesh.bbox.bbox1 = mesh.bbox.bbox0;
  for (int i = 0; i < hdr[0]; i++)
    for (int d = 0; d < 3; d++) {
      const float v = vbo[3 * i + d];
      if (v < mesh.bbox.bbox0[d]) mesh.bbox.bbox0[d] = v;
      if (v > mesh.bbox.bbox1[d]) mesh.bbox.bbox1[d] = v;
    }
  bmp_texture no_texture;  // -> the reference's ColorShader paints green (voxelization.cu:100-102)
  world::Scene scene;
  scene.addMesh(mesh, no_texture);
  scene.voxelizeMeshes(true);
  BoundingBox any;
  SVO svo = scene.svo(any);
  const int n_nodes = scene.tree()->nodeCount();
  std::vector<uint32_t> pool((size_t)n_nodes * 2);
  cudaMemcpy(pool.data(), svo.data, pool.size() * 4, cudaMemcpyDeviceToHost);
  char path[4096];
  snprintf(path, sizeof(path), "%s.pool", out_prefix);
  FILE* o = fopen(path, "wb");
  if (!o) { perror(path); return 2; }
  fwrite(&n_nodes, 4, 1, o);
  fwrite(&svo.center.x, 4, 3, o);
  fwrite(&svo.size, 4, 1, o);
  fwrite(pool.data(), 4, pool.size(), o);
  fclose(o);
  printf("osl_main: mesh %d vertices %d triangles -> %d voxels, %d nodes (center %.6f %.6f %.6f, half %.6f)\n", hdr[0],
         hdr[1], scene.voxel_grid().size, n_nodes, svo.center.x, svo.center.y, svo.center.z, svo.size);
  return 0;
}

int main(int argc, char** argv) {
  // the frame pipeline's streams are served best by 3 hardware work queues; must precede
  // the first CUDA call, an explicit setting in the environment wins
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "3", 0);
  if (argc == 4 && !strcmp(argv[1], "mesh")) return mesh_main(argv[2], argv[3]);
  if (argc < 3) {
    fprintf(stderr, "usage: %s frames.bin out_prefix [fused|track|slam]\n", argv[0]);
    return 2;
  }
  const bool fused = argc > 3 && !strcmp(argv[3], "fused");
  const bool slam = argc > 3 && !strcmp(argv[3], "slam");  // track + the pose stays on the device (fused frames)
  const bool track = slam || (argc > 3 && !strcmp(argv[3], "track"));
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int hdr[3];
  float focal[2];
  if (!read_exact(f, hdr, sizeof(hdr)) || !read_exact(f, focal, sizeof(focal))) return 2;
  const int W = hdr[0], H = hdr[1], n_frames = hdr[2];
  const int num_points = W * H;

  // init() (main.cpp:86-150) minus GLFW / OpenNI
  RawFrame frame(W, H);                                   // OpenNIDevice::raw_frame_
  glm::vec3* points_ = nullptr;                           // main.cpp:139
  cudaMalloc((void**)&points_, (size_t)num_points * sizeof(glm::vec3));
  world::Scene* scene_ = new world::Scene();
  rendering::CUDARenderer* cuda_renderer_ = new rendering::CUDARenderer(W, H);
  Camera camera;                                          // GLFWCameraController::camera(): fov = 45
  camera.view[0].x = -1.0f;                               // look down +z like the depth camera (the reference's
  camera.view[2].z = -1.0f;                               // renderer looks down -z for view = identity)
  const glm::vec2 focal_length(focal[0], focal[1]);

  sensor::RGBDCamera* camera_estimation_ = track ? new sensor::RGBDCamera(W, H, focal_length, true) : nullptr;
  std::vector<float> est_poses;

  std::vector<uint16_t> h_depth((size_t)num_points);
  std::vector<uint8_t> h_rgb((size_t)num_points * 3);
  BoundingBox cloud_bbox;
  for (int k = 0; k < n_frames; k++) {
    glm::mat4 pose;
    if (!read_exact(f, &pose[0].x, 64) || !read_exact(f, h_depth.data(), h_depth.size() * 2) ||
        !read_exact(f, h_rgb.data(), h_rgb.size()))
      return 2;
    // OpenNIDevice::readFrame (openni_device.cpp:122,144)
    cudaMemcpy(frame.depth, h_depth.data(), h_depth.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(frame.color, h_rgb.data(), h_rgb.size(), cudaMemcpyHostToDevice);

    if (track) {  // main.cpp:35 + :40
      frame.timestamp = k + 1;
      camera_estimation_->update(&frame);
      pose = camera_estimation_->pose();
      est_poses.insert(est_poses.end(), &pose[0].x, &pose[0].x + 16);
    }

    // main.cpp:38-44
    sensor::generateVertexMap(frame.depth, points_, W, H, focal_length, make_int2(W, H));
    sensor::transformVertexMap(points_, pose, W * H);
    cudaDeviceSynchronize();
    cloud_bbox = BoundingBox();
    sensor::computePointCloudBoundingBox(points_, num_points, cloud_bbox);
    if (slam && scene_->tree()) {
      // tracker -> integration of the same frame with no host round trip (the pose above was only read for the log)
      scene_->tree()->addDepthFrame(frame.depth, frame.color, W, H, focal_length, *camera_estimation_);
    } else if (fused && scene_->tree()) {
      // the one-call fast path for every frame after the one that creates the tree
      scene_->tree()->addDepthFrame(frame.depth, frame.color, W, H, focal_length, pose);
    } else {
      scene_->addPointCloudToOctree(glm::vec3(pose[3].x, pose[3].y, pose[3].z), points_, frame.color, num_points,
                                    cloud_bbox);
    }
  }
  fclose(f);

  // main.cpp:56-58
  SVO svo = scene_->svo(cloud_bbox);
  cuda_renderer_->coneTraceSVO(svo, camera, glm::vec3(0.0f));

  const int n_nodes = scene_->tree()->nodeCount();
  std::vector<uint32_t> pool((size_t)n_nodes * 2);
  cudaMemcpy(pool.data(), svo.data, pool.size() * 4, cudaMemcpyDeviceToHost);
  std::vector<uchar4> img((size_t)num_points);
  cuda_renderer_->download(img.data());

  char path[4096];
  snprintf(path, sizeof(path), "%s.pool", argv[2]);
  FILE* o = fopen(path, "wb");
  if (!o) { perror(path); return 2; }
  fwrite(&n_nodes, 4, 1, o);
  fwrite(&svo.center.x, 4, 3, o);
  fwrite(&svo.size, 4, 1, o);
  fwrite(pool.data(), 4, pool.size(), o);
  fclose(o);
  snprintf(path, sizeof(path), "%s.rgba", argv[2]);
  o = fopen(path, "wb");
  if (!o) { perror(path); return 2; }
  fwrite(img.data(), 4, img.size(), o);
  fclose(o);
  if (track) {
    snprintf(path, sizeof(path), "%s.poses", argv[2]);
    o = fopen(path, "wb");
    if (!o) { perror(path); return 2; }
    fwrite(est_poses.data(), 4, est_poses.size(), o);
    fclose(o);
    delete camera_estimation_;
  }
  printf("osl_main: %d frames %dx%d -> %d nodes (center %.6f %.6f %.6f, half %.6f)\n", n_frames, W, H, n_nodes,
         svo.center.x, svo.center.y, svo.center.z, svo.size);

  delete cuda_renderer_;
  delete scene_;
  cudaFree(points_);
  return 0;
}
