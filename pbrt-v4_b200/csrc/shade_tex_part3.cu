// shade_tex_part3.cu -- compilation 4 of 4 of shade_tex.cu (two TEX = true shading kernels per object file, see there)
#define BPT_TEX_PART 3
#include "shade_tex.cu"
