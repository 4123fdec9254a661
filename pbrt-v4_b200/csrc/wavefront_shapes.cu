// wavefront_shapes.cu -- the second compilation of wavefront.cu: the same wavefront kernels with the quadric leaf test of the traversal
// (bvh.cuh LeafIntersect<true>), the quadric surface interactions and the quadric area lights compiled in (lights.h, quadrics.h).
// Everything lands in namespace bpt::tu_shapes; the engine launches this set for scenes that contain Sphere / Disk / Cylinder shapes
// and the plain set otherwise, so the triangle-only kernels keep their registers and instruction footprint.
#define BPT_WITH_SHAPES 1
#include "wavefront.cu"
