// Test tool: Triangle::InteractionFromIntersection as restated in csrc/shapes.h (TriangleInteraction: hit point with its error
// bounds, geometric and shading normals, uv, shading dpdu; per-vertex normals / uv, reversed orientation, degenerate uv) compiled
// for the CPU against golden vectors written by the reference itself (oracle/ref_dump.cpp "triintr{0..3}_*").
//   g++ -O2 -ffp-contract=off -mfma -std=c++17 -I/usr/local/cuda/include -Ipbrt-v4_b200/csrc tests/tools/triintr_host_check.cpp -o /tmp/tricheck
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "shapes.h"
using namespace bpt;
namespace bpt { uint32_t h_sobol32[104]; }
static bool eq(float a, float b) { return f2b(a) == f2b(b) || (a != a && b != b); }
int main() {
    std::map<std::string, std::vector<float>> g;
    FILE *f = fopen("tests/golden/ref_vectors.bin", "rb"); char magic[8];
    if (!f || fread(magic, 1, 8, f) != 8) return 3;
    while (true) { char nm[33]; uint32_t dt; uint64_t n; if (fread(nm, 1, 32, f) != 32) break; nm[32] = 0; if (fread(&dt, 4, 1, f) != 1 || fread(&n, 8, 1, f) != 1) break;
        std::vector<float> v(n); if (fread(v.data(), 4, n, f) != n) break; g[nm] = v; }
    fclose(f);
    int total = 0;
    for (int cfg = 0; cfg < 4; ++cfg) {
        char nm[64];
        snprintf(nm, 64, "triintr%d_mesh", cfg); const std::vector<float> &md = g[nm];
        snprintf(nm, 64, "triintr%d_in", cfg); const std::vector<float> &in = g[nm];
        snprintf(nm, 64, "triintr%d_out", cfg); const std::vector<float> &out = g[nm];
        if (md.empty()) { printf("missing golden\n"); return 3; }
        TriAttribs base; std::memset(&base, 0, sizeof(base));
        base.hasN = md[1] != 0; base.hasUV = md[2] != 0; base.flipNormal = md[3] != 0;
        auto V = [&](int vert, int k) { return md[4 + 8 * (size_t)vert + k]; };
        int n = (int)in.size() / 4, bad = 0;
        for (int q = 0; q < n; ++q) {
            const float *a = &in[4 * q], *o = &out[17 * q];
            int tri = (int)a[0];
            Pt3 p[3]; TriAttribs at = base;
            Nrm3 nn[3]; Vec2 uv[3];
            for (int k = 0; k < 3; ++k) {
                int v = 3 * tri + k;
                p[k] = Pt3(V(v, 0), V(v, 1), V(v, 2)); nn[k] = Nrm3(V(v, 3), V(v, 4), V(v, 5)); uv[k] = Vec2(V(v, 6), V(v, 7));
            }
            at.n0 = nn[0]; at.n1 = nn[1]; at.n2 = nn[2]; at.uv0 = uv[0]; at.uv1 = uv[1]; at.uv2 = uv[2];
            TriHit th{a[1], a[2], a[3], 1.f};
            SurfaceIntr si = TriangleInteraction(p[0], p[1], p[2], at, th);
            const float mine[17] = {si.pi.x.low, si.pi.x.high, si.pi.y.low, si.pi.y.high, si.pi.z.low, si.pi.z.high, si.n.x, si.n.y, si.n.z,
                                    si.ns.x, si.ns.y, si.ns.z, si.uv.x, si.uv.y, si.dpdus.x, si.dpdus.y, si.dpdus.z};
            bool ok = true;
            for (int k = 0; k < 17; ++k) ok = ok && eq(mine[k], o[k]);
            if (!ok && bad++ < 3) {
                printf("cfg %d query %d tri %d:", cfg, q, tri);
                for (int k = 0; k < 17; ++k) if (!eq(mine[k], o[k])) printf(" [%d] %g vs %g", k, mine[k], o[k]);
                printf("\n");
            }
        }
        printf("config %d (normals %d, uv %d, flip %d): %d mismatches of %d\n", cfg, (int)base.hasN, (int)base.hasUV, (int)base.flipNormal, bad, n);
        total += bad;
    }
    return total ? 2 : 0;
}
