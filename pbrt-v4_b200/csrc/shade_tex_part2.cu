// shade_tex_part2.cu -- compilation 3 of 4 of shade_tex.cu (two TEX = true shading kernels per object file, see there)
#define BPT_TEX_PART 2
#include "shade_tex.cu"
