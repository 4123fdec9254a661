// shade_tex_layered.cu -- the second compilation of shade_tex.cu: k_shade<CoatedConductor / CoatedDiffuse, TEX = true> (the two kernels
// that take ptxas longest), in an object file of their own so that they compile beside the single-layer materials.
#define BPT_TEX_LAYERED 1
#include "shade_tex.cu"
