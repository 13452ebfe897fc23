#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "flux_bvh.cuh"
bool build_bvh4(const double *sph, const uint32_t *sph_meta, uint32_t ns, const double *tri, const uint32_t *tri_meta,
                const double *tri_v1, const double *tri_v2, uint32_t nt, BvhBuild &out, std::string &err);
int main() {
    const uint32_t nt = 300000, ns = 3;
    std::mt19937_64 g(1);
    std::uniform_real_distribution<double> U(-50, 50), S(0, 0.3);
    std::vector<double> tri((size_t)TRI_FIELDS * nt), v1(3 * (size_t)nt), v2(3 * (size_t)nt), sph((size_t)SPH_FIELDS * ns, 0.0);
    std::vector<uint32_t> tm(2 * (size_t)nt, 0), sm(2 * ns, 0);
    for (uint32_t i = 0; i < nt; i++)
        for (int k = 0; k < 3; k++) {
            double a = U(g), e1 = S(g), e2 = S(g);
            tri[(size_t)(TRI_V0X + k) * nt + i] = a; tri[(size_t)(TRI_E1X + k) * nt + i] = e1; tri[(size_t)(TRI_E2X + k) * nt + i] = e2;
            v1[3 * (size_t)i + k] = a + e1; v2[3 * (size_t)i + k] = a + e2;
        }
    for (uint32_t i = 0; i < ns; i++) {
        double c[3] = {U(g), U(g), U(g)}, r = i == 0 ? 5000.0 : 1.0;
        sph[(size_t)SPH_CX * ns + i] = c[0]; sph[(size_t)SPH_CY * ns + i] = c[1]; sph[(size_t)SPH_CZ * ns + i] = c[2];
        sph[(size_t)SPH_C0X * ns + i] = c[0] - r; sph[(size_t)SPH_C1X * ns + i] = c[0] + r;
        sph[(size_t)SPH_C0Y * ns + i] = c[1] - r; sph[(size_t)SPH_C1Y * ns + i] = c[1] + r;
        sph[(size_t)SPH_C0Z * ns + i] = c[2] - r; sph[(size_t)SPH_C1Z * ns + i] = c[2] + r;
        sph[(size_t)SPH_R * ns + i] = r; sph[(size_t)SPH_RR * ns + i] = r * r;
    }
    BvhBuild a, b;
    std::string err;
    if (!build_bvh4(sph.data(), sm.data(), ns, tri.data(), tm.data(), v1.data(), v2.data(), nt, a, err)) { std::printf("fail %s\n", err.c_str()); return 1; }
    if (!build_bvh4(sph.data(), sm.data(), ns, tri.data(), tm.data(), v1.data(), v2.data(), nt, b, err)) { std::printf("fail %s\n", err.c_str()); return 1; }
    const bool same = a.nodes.size() == b.nodes.size() && !std::memcmp(a.nodes.data(), b.nodes.data(), a.nodes.size() * sizeof(BvhNode4)) &&
                      a.prims == b.prims && a.linear == b.linear;
    std::printf("nodes %zu prims %zu linear %zu depth %u same %d\n", a.nodes.size(), a.prims.size(), a.linear.size(), a.depth, (int)same);
    return same ? 0 : 2;
}
