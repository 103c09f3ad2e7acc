/* Empty stand-in: the reference's image_kernels.h includes <OpenNI.h> ("TODO: Why is this needed?",
 * include/octree_slam/sensor/image_kernels.h:11) but uses nothing from it. */
