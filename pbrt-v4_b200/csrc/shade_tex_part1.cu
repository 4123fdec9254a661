// shade_tex_part1.cu -- compilation 2 of 4 of shade_tex.cu (two TEX = true shading kernels per object file, see there)
#define BPT_TEX_PART 1
#include "shade_tex.cu"
