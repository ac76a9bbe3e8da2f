// mesh.cpp -- uniform-refinement mesh generator (host code, no CUDA).
//
// Produces the particle set (x, y, z, area, isActive) of an lpm-v2 PolyMesh2d
// refined uniformly to `initNest`, in the reference's exact particle insertion
// order, so that synthetic inputs for the direct-sum kernels have the same
// target order, active-source list and panel areas the reference would hand
// to its solvers.  Written from the behaviour of:
//   src/PolyMesh2d.f90:135-195   (New: seed, then divide every face of the previous level, in order)
//   src/PolyMesh2d.f90:795-939   (initializeMeshFromSeed: vertices first, then face centres;
//                                 icosTri centres recomputed with SphereTriCenter :875-879;
//                                 beta-plane seed affine map :880-884)
//   src/Faces.f90:701-858        (DivideTriFace: walk the 3 parent edges, split un-split ones,
//                                 3 new centre particles, child 4 keeps the parent's centre)
//   src/Faces.f90:529-689        (DivideQuadFace: 4 new centres, parent centre becomes passive)
//   src/Edges.f90:567-621        (divideLinearEdge: midpoint particle, two child edges)
//   src/Faces.f90:917-975        (QuadFaceArea / TriFaceArea: fan of sub-triangles about the centre)
//   src/SphereGeometry.f90:168-188,280-334,347-367 (arc length, midpoint, centres, triangle area)
//   src/PlaneGeometry.f90:74-113 (centroids, triangle area)
// The five seed tables below are the data of the reference's *Seed.dat files
// (vertex coordinates, edge orig/dest/left/right, face vertices/edges, 0-based
// as stored there); they are input data, not code.
//
//
// Pinned: tests/test_refsrc_golden.py::test_mesh_generator_bits compares this generator, bit for bit, with the
// reference's own PolyMesh2d New executed from its source text (and its own *Seed.dat) by the test infrastructure's
// Fortran-subset interpreter --
// coordinates, areas, particle order, active flags, tree sizes, MaxEdgeLength and leaf-face connectivity for all five
// seeds at levels 0-3.
//
// Only uniform refinement is implemented (no AMR); remeshing stays in the
// reference Fortran (out of scope).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "lpm_mesh.h"

namespace {

struct Seed {
    int nParticles, nEdges, nFaces, nVerts, vertsPerFace;
    bool sphere;
    const double* xyz;      // nVerts*3 (sphere) or nParticles*2 (plane); see initialize()
    const int* edges;       // nEdges*4: orig dest left right (0-based, -1 = boundary)
    const int* faceVerts;   // nFaces*vertsPerFace
    const int* faceEdges;   // nFaces*vertsPerFace
};

// ---- icosTriSeed.dat --------------------------------------------------------
const double kIcosXYZ[12 * 3] = {
    0.0, 0.0, 1.0,
    0.723606797749978969640917366873, 0.525731112119133606025669084848, 0.447213595499957939281834733746,
    -0.276393202250021030359082633126, 0.850650808352039932181540497063, 0.447213595499957939281834733746,
    -0.894427190999915878563669467492, 0.0, 0.447213595499957939281834733746,
    -0.276393202250021030359082633127, -0.850650808352039932181540497063, 0.447213595499957939281834733746,
    0.723606797749978969640917366873, -0.525731112119133606025669084848, 0.447213595499957939281834733746,
    0.894427190999915878563669467492, 0.0, -0.447213595499957939281834733746,
    0.276393202250021030359082633127, 0.850650808352039932181540497063, -0.447213595499957939281834733746,
    -0.723606797749978969640917366873, 0.525731112119133606025669084848, -0.447213595499957939281834733746,
    -0.723606797749978969640917366873, -0.525731112119133606025669084848, -0.447213595499957939281834733746,
    0.276393202250021030359082633127, -0.850650808352039932181540497063, -0.447213595499957939281834733746,
    0.0, 0.0, -1.0};
const int kIcosEdges[30 * 4] = {
    0, 1, 0, 4,    1, 2, 0, 6,    2, 0, 0, 1,    2, 3, 1, 8,    0, 3, 2, 1,    3, 4, 2, 10,
    4, 0, 2, 3,    4, 5, 3, 12,   5, 0, 3, 4,    5, 1, 4, 14,   1, 6, 5, 14,   6, 7, 5, 15,
    7, 1, 5, 6,    7, 2, 6, 7,    7, 8, 7, 16,   8, 2, 7, 8,    8, 3, 8, 9,    8, 9, 9, 17,
    3, 9, 10, 9,   9, 4, 10, 11,  9, 10, 11, 18, 10, 4, 11, 12, 10, 5, 12, 13, 10, 6, 13, 19,
    6, 5, 13, 14,  11, 6, 19, 15, 11, 7, 15, 16, 8, 11, 17, 16, 11, 9, 17, 18, 10, 11, 19, 18};
const int kIcosFaceVerts[20 * 3] = {
    0, 1, 2,   0, 2, 3,   0, 3, 4,   0, 4, 5,   0, 5, 1,   1, 6, 7,   7, 2, 1,   2, 7, 8,   8, 3, 2,   3, 8, 9,
    9, 4, 3,   4, 9, 10,  10, 5, 4,  5, 10, 6,  6, 1, 5,   11, 7, 6,  11, 8, 7,  11, 9, 8,  11, 10, 9, 11, 6, 10};
const int kIcosFaceEdges[20 * 3] = {
    0, 1, 2,     2, 3, 4,     4, 5, 6,     6, 7, 8,     8, 9, 0,     10, 11, 12,  13, 1, 12,
    13, 14, 15,  16, 3, 15,   16, 17, 18,  19, 5, 18,   19, 20, 21,  22, 7, 21,   22, 23, 24,
    10, 9, 24,   26, 11, 25,  27, 14, 26,  28, 17, 27,  29, 20, 28,  25, 23, 29};

// ---- cubedSphereSeed.dat ----------------------------------------------------
#define CS 0.577350269189626
const double kCubeXYZ[14 * 3] = {
    CS, -CS, CS,   CS, -CS, -CS,  CS, CS, -CS,   CS, CS, CS,    -CS, CS, -CS,  -CS, CS, CS,  -CS, -CS, -CS,
    -CS, -CS, CS,  1.0, 0.0, 0.0, 0.0, 1.0, 0.0, -1.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, -1.0};
#undef CS
const int kCubeEdges[12 * 4] = {
    0, 1, 0, 3,  1, 2, 0, 6,  2, 3, 0, 1,  3, 0, 0, 4,  2, 4, 1, 5,  4, 5, 1, 2,
    5, 3, 1, 4,  4, 6, 2, 5,  6, 7, 2, 3,  7, 5, 2, 4,  6, 1, 3, 5,  0, 7, 3, 4};
const int kCubeFaceVerts[6 * 4] = {0, 1, 2, 3,  3, 2, 4, 5,  5, 4, 6, 7,  7, 6, 1, 0,  7, 0, 3, 5,  1, 6, 4, 2};
const int kCubeFaceEdges[6 * 4] = {0, 1, 2, 3,  2, 4, 5, 6,  5, 7, 8, 9,  8, 10, 0, 11,  11, 3, 6, 9,  10, 7, 4, 1};

// ---- quadRectSeed.dat / betaPlaneSeed.dat (same coordinates and faces) ------
const double kQuadXY[13 * 2] = {-1.0, 1.0,  -1.0, 0.0,  -1.0, -1.0, 0.0, -1.0, 1.0, -1.0, 1.0, 0.0, 1.0, 1.0,
                                0.0, 1.0,   0.0, 0.0,   -0.5, 0.5,  -0.5, -0.5, 0.5, -0.5, 0.5, 0.5};
const int kQuadEdges[12 * 4] = {0, 1, 0, -1, 1, 2, 1, -1, 2, 3, 1, -1, 3, 4, 2, -1, 4, 5, 2, -1, 5, 6, 3, -1,
                                6, 7, 3, -1, 7, 0, 0, -1, 1, 8, 0, 1,  8, 5, 3, 2,  3, 8, 1, 2,  8, 7, 0, 3};
const int kBetaEdges[12 * 4] = {0, 1, 0, 3,  1, 2, 1, 2,  2, 3, 1, -1, 3, 4, 2, -1, 4, 5, 2, 2,  5, 6, 3, 0,
                                6, 7, 3, -1, 7, 0, 0, -1, 1, 8, 0, 1,  8, 5, 3, 2,  3, 8, 1, 2,  8, 7, 0, 3};
const int kQuadFaceVerts[4 * 4] = {0, 1, 8, 7,  1, 2, 3, 8,  8, 3, 4, 5,  7, 8, 5, 6};
const int kQuadFaceEdges[4 * 4] = {0, 8, 11, 7,  1, 2, 10, 8,  10, 3, 4, 9,  11, 9, 5, 6};

// ---- triHexSeed.dat ---------------------------------------------------------
const double kHexXY[13 * 2] = {
    0.0, 0.0,  0.5, 0.866025403784438597,  -0.5, 0.866025403784438597,  -1.0, 0.0,
    -0.5, -0.866025403784438597,  0.5, -0.866025403784438597,  1.0, 0.0,
    0.0, 0.577350269189625731,  -0.5, 0.288675134594812810,  -0.5, -0.288675134594812810,
    0.0, -0.577350269189625731, 0.5, -0.288675134594812810,  0.5, 0.288675134594812810};
const int kHexEdges[12 * 4] = {0, 1, 0, 5,  1, 2, 0, -1, 2, 3, 1, -1, 3, 4, 2, -1, 4, 5, 3, -1, 5, 6, 4, -1,
                               6, 1, 5, -1, 2, 0, 0, 1,  3, 0, 1, 2,  4, 0, 2, 3,  0, 5, 4, 3,  0, 6, 5, 4};
const int kHexFaceVerts[6 * 3] = {0, 1, 2,  2, 3, 0,  4, 0, 3,  0, 4, 5,  5, 6, 0,  1, 0, 6};
const int kHexFaceEdges[6 * 3] = {0, 1, 7,  2, 8, 7,  9, 8, 3,  9, 4, 10,  5, 11, 10,  0, 11, 6};

const Seed kSeedIcos = {32, 30, 20, 12, 3, true, kIcosXYZ, kIcosEdges, kIcosFaceVerts, kIcosFaceEdges};
const Seed kSeedCube = {14, 12, 6, 8, 4, true, kCubeXYZ, kCubeEdges, kCubeFaceVerts, kCubeFaceEdges};
const Seed kSeedQuad = {13, 12, 4, 9, 4, false, kQuadXY, kQuadEdges, kQuadFaceVerts, kQuadFaceEdges};
const Seed kSeedBeta = {13, 12, 4, 9, 4, false, kQuadXY, kBetaEdges, kQuadFaceVerts, kQuadFaceEdges};
const Seed kSeedHex = {13, 12, 6, 7, 3, false, kHexXY, kHexEdges, kHexFaceVerts, kHexFaceEdges};

// src/TypeDefs.f90:74 -- the module-global SphereRadius is never assigned by
// the reference; every sphere projection multiplies by this 1.0.
const double kSphereRadiusGlobal = 1.0;

struct V3 { double x, y, z; };

struct Mesh {
    int seedKind = 0;
    int vpf = 3;        // vertices per face
    bool sphere = true;
    // particles
    std::vector<double> x, y, z, area;
    std::vector<int32_t> active;
    // edges (0-based indices; face index -1 = none)
    std::vector<int32_t> eOrig, eDest, eLeft, eRight, eChild1, eChild2;
    std::vector<uint8_t> eHasKids;
    // faces
    std::vector<int32_t> fVerts, fEdges, fCenter;
    std::vector<uint8_t> fHasKids;
    int64_t nFaces() const { return (int64_t)fCenter.size(); }
    int64_t nEdges() const { return (int64_t)eOrig.size(); }
    int64_t nParticles() const { return (int64_t)x.size(); }

    V3 P(int32_t i) const { return V3{x[i], y[i], z[i]}; }

    int32_t insertParticle(V3 p) {       // src/Particles.f90:496-516 (inserted passive)
        x.push_back(p.x); y.push_back(p.y); z.push_back(sphere ? p.z : 0.0);
        area.push_back(0.0); active.push_back(0);
        return (int32_t)x.size() - 1;
    }
    int32_t insertEdge(int32_t o, int32_t d, int32_t l, int32_t r) {   // src/Edges.f90:211-233
        eOrig.push_back(o); eDest.push_back(d); eLeft.push_back(l); eRight.push_back(r);
        eChild1.push_back(-1); eChild2.push_back(-1); eHasKids.push_back(0);
        return (int32_t)eOrig.size() - 1;
    }

    // ---- geometry ----------------------------------------------------------
    static double arcLength(V3 a, V3 b) {          // src/SphereGeometry.f90:168-188
        double c0 = a.y * b.z - b.y * a.z, c1 = b.x * a.z - a.x * b.z, c2 = a.x * b.y - b.x * a.y;
        double crossNorm = std::sqrt(c0 * c0 + c1 * c1 + c2 * c2);
        double dotProd = a.x * b.x + a.y * b.y + a.z * b.z;
        return std::atan2(crossNorm, dotProd);
    }
    static V3 project(V3 s) {                      // "/sqrt(sum(v*v))*SphereRadius"
        double nrm = std::sqrt(s.x * s.x + s.y * s.y + s.z * s.z);
        return V3{s.x / nrm * kSphereRadiusGlobal, s.y / nrm * kSphereRadiusGlobal, s.z / nrm * kSphereRadiusGlobal};
    }
    static V3 sphereMidpoint(V3 a, V3 b) {         // src/SphereGeometry.f90:280-288
        return project(V3{(a.x + b.x) / 2.0, (a.y + b.y) / 2.0, (a.z + b.z) / 2.0});
    }
    static V3 sphereTriCenter(V3 a, V3 b, V3 c) {  // src/SphereGeometry.f90:302-311
        return project(V3{(a.x + b.x + c.x) / 3.0, (a.y + b.y + c.y) / 3.0, (a.z + b.z + c.z) / 3.0});
    }
    static V3 sphereQuadCenter(V3 a, V3 b, V3 c, V3 d) {   // src/SphereGeometry.f90:326-335
        return project(V3{(a.x + b.x + c.x + d.x) / 4.0, (a.y + b.y + c.y + d.y) / 4.0, (a.z + b.z + c.z + d.z) / 4.0});
    }
    static double sphereTriArea(V3 a, V3 b, V3 c) {        // src/SphereGeometry.f90:347-367
        double s1 = arcLength(a, b), s2 = arcLength(b, c), s3 = arcLength(c, a);
        double hp = (s1 + s2 + s3) / 2.0;
        double zz = std::tan(hp / 2.0) * std::tan((hp - s1) / 2.0) * std::tan((hp - s2) / 2.0) * std::tan((hp - s3) / 2.0);
        return 4.0 * std::atan2(std::sqrt(zz), 1.0) * kSphereRadiusGlobal * kSphereRadiusGlobal;
    }
    static double planeTriArea(V3 a, V3 b, V3 c) {         // src/PlaneGeometry.f90:109-113
        return 0.5 * std::fabs(-b.x * a.y + c.x * a.y + a.x * b.y - c.x * b.y - a.x * c.y + b.x * c.y);
    }
    double faceArea(int64_t f) const {                     // src/Faces.f90:917-975
        double A = 0.0;
        V3 c = P(fCenter[f]);
        for (int i = 0; i < vpf; ++i) {
            V3 v1 = P(fVerts[f * vpf + i]), v2 = P(fVerts[f * vpf + (i + 1) % vpf]);
            A = A + (sphere ? sphereTriArea(v1, c, v2) : planeTriArea(v1, c, v2));
        }
        return A;
    }

    // ---- src/Edges.f90:567-621 divideLinearEdge -----------------------------
    void divideEdge(int32_t e) {
        V3 v0 = P(eOrig[e]), v1 = P(eDest[e]);
        V3 mid = sphere ? sphereMidpoint(v0, v1) : V3{0.5 * (v0.x + v1.x), 0.5 * (v0.y + v1.y), 0.0};
        int32_t p = insertParticle(mid);
        int32_t c1 = insertEdge(eOrig[e], p, eLeft[e], eRight[e]);
        // The reference assigns rightFace(N+1) twice and never rightFace(N+2)
        // (Edges.f90:601-605); child 2's right face is filled in by the face
        // division that follows.  Particle order does not depend on it.
        int32_t c2 = insertEdge(p, eDest[e], eLeft[e], -1);
        eHasKids[e] = 1; eChild1[e] = c1; eChild2[e] = c2;
    }

    // ---- src/Faces.f90:701-858 / 529-689 ------------------------------------
    void divideFace(int64_t f) {
        const int K = vpf;
        const int32_t nF = (int32_t)nFaces();
        int32_t nv[4][4], ne[4][4];     // [slot][child]
        std::memset(nv, 0xff, sizeof(nv)); std::memset(ne, 0xff, sizeof(ne));
        for (int i = 0; i < K; ++i) nv[i][i] = fVerts[f * K + i];
        for (int i = 0; i < K; ++i) {
            const int in = (i + 1) % K;
            int32_t pe = fEdges[f * K + i];
            if (!eHasKids[pe]) divideEdge(pe);
            int32_t c1 = eChild1[pe], c2 = eChild2[pe];
            if ((int32_t)f == eLeft[pe]) {                // positiveEdge, src/Edges.f90:678-685
                ne[i][i] = c1;  eLeft[c1] = nF + i;
                ne[i][in] = c2; eLeft[c2] = nF + in;
            } else {
                ne[i][i] = c2;  eRight[c2] = nF + i;
                ne[i][in] = c1; eRight[c1] = nF + in;
            }
            nv[i][in] = eDest[c1];
            nv[in][i] = eDest[c1];
        }
        if (K == 3) {
            nv[0][3] = nv[2][1]; nv[1][3] = nv[0][2]; nv[2][3] = nv[1][0];
            int32_t e;
            e = insertEdge(nv[0][3], nv[1][3], nF + 3, nF + 2); ne[0][3] = e; ne[0][2] = e;
            e = insertEdge(nv[1][3], nv[2][3], nF + 3, nF + 0); ne[1][3] = e; ne[1][0] = e;
            e = insertEdge(nv[2][3], nv[0][3], nF + 3, nF + 1); ne[2][3] = e; ne[2][1] = e;
        } else {
            for (int i = 0; i < 4; ++i) nv[(i + 2) % 4][i] = fCenter[f];
            int32_t e;
            e = insertEdge(nv[1][0], nv[2][0], nF + 0, nF + 1); ne[1][0] = e; ne[3][1] = e;
            e = insertEdge(nv[0][2], nv[3][2], nF + 3, nF + 2); ne[3][2] = e; ne[1][3] = e;
            e = insertEdge(nv[2][1], nv[3][1], nF + 1, nF + 2); ne[2][1] = e; ne[0][2] = e;
            e = insertEdge(nv[1][3], nv[0][3], nF + 0, nF + 3); ne[0][3] = e; ne[2][0] = e;
        }
        // centres are computed for all children before any is inserted
        V3 ctr[4];
        for (int c = 0; c < 4; ++c) {
            V3 v[4];
            for (int j = 0; j < K; ++j) v[j] = P(nv[j][c]);
            if (K == 3)
                ctr[c] = sphere ? sphereTriCenter(v[0], v[1], v[2])
                                : V3{(v[0].x + v[1].x + v[2].x) / 3.0, (v[0].y + v[1].y + v[2].y) / 3.0, 0.0};
            else
                ctr[c] = sphere ? sphereQuadCenter(v[0], v[1], v[2], v[3])
                                : V3{0.25 * (v[0].x + v[1].x + v[2].x + v[3].x), 0.25 * (v[0].y + v[1].y + v[2].y + v[3].y), 0.0};
        }
        const int nNewCenters = (K == 3) ? 3 : 4;
        for (int c = 0; c < 4; ++c) {
            int32_t cp;
            if (c < nNewCenters) { cp = insertParticle(ctr[c]); active[cp] = 1; }
            else cp = fCenter[f];                          // tri child 4 reuses the parent's centre
            fCenter.push_back(cp);
            for (int j = 0; j < K; ++j) { fVerts.push_back(nv[j][c]); fEdges.push_back(ne[j][c]); }
            fHasKids.push_back(0);
            area[cp] = faceArea(nF + c);
        }
        if (K == 4) { active[fCenter[f]] = 0; area[fCenter[f]] = 0.0; }   // src/Faces.f90:684-685
        fHasKids[f] = 1;
    }

    void initialize(const Seed& s, int kind, double ampFactor) {
        seedKind = kind; vpf = s.vertsPerFace; sphere = s.sphere;
        std::vector<V3> pts(s.nParticles);
        if (s.sphere) {
            int nRead = (kind == LPM_ICOS_TRI_SPHERE_SEED) ? 12 : s.nParticles;
            for (int i = 0; i < nRead; ++i) pts[i] = V3{s.xyz[3 * i], s.xyz[3 * i + 1], s.xyz[3 * i + 2]};
            if (kind == LPM_ICOS_TRI_SPHERE_SEED)
                for (int i = 0; i < 20; ++i)
                    pts[12 + i] = sphereTriCenter(pts[s.faceVerts[3 * i]], pts[s.faceVerts[3 * i + 1]], pts[s.faceVerts[3 * i + 2]]);
        } else {
            for (int i = 0; i < s.nParticles; ++i) pts[i] = V3{s.xyz[2 * i], s.xyz[2 * i + 1], 0.0};
            if (kind == LPM_BETA_PLANE_SEED)
                for (int i = 0; i < 13; ++i) { pts[i].x = 0.5 * pts[i].x + 0.5; pts[i].y = 0.5 * pts[i].y; }
        }
        for (auto& p : pts) { p.x = ampFactor * p.x; p.y = ampFactor * p.y; p.z = ampFactor * p.z; }
        for (int i = 0; i < s.nParticles; ++i) insertParticle(pts[i]);
        for (int i = 0; i < s.nEdges; ++i)
            insertEdge(s.edges[4 * i], s.edges[4 * i + 1], s.edges[4 * i + 2], s.edges[4 * i + 3]);
        for (int i = 0; i < s.nFaces; ++i) {
            fCenter.push_back(s.nVerts + i);
            for (int j = 0; j < vpf; ++j) { fVerts.push_back(s.faceVerts[vpf * i + j]); fEdges.push_back(s.faceEdges[vpf * i + j]); }
            fHasKids.push_back(0);
            active[s.nVerts + i] = 1;
        }
        for (int i = 0; i < s.nFaces; ++i) area[fCenter[i]] = faceArea(i);
    }

    void refine(int initNest) {                    // src/PolyMesh2d.f90:178-190
        int64_t start = 0;
        for (int lev = 0; lev < initNest; ++lev) {
            int64_t nOld = nFaces();
            for (int64_t j = start; j < nOld; ++j) divideFace(j);
            start = nOld;
        }
    }

    double maxEdgeLength() const {                 // src/Edges.f90:260-297
        double m = 0.0;
        for (int64_t e = 0; e < nEdges(); ++e) {
            if (eHasKids[e]) continue;
            V3 a = P(eOrig[e]), b = P(eDest[e]);
            double len = sphere ? arcLength(a, b) * kSphereRadiusGlobal
                                : std::sqrt((b.x - a.x) * (b.x - a.x) + (b.y - a.y) * (b.y - a.y) + (b.z - a.z) * (b.z - a.z));
            if (len > m) m = len;
        }
        return m;
    }
};

}  // namespace

struct lpm_mesh { Mesh m; };

extern "C" {

int lpm_mesh_create(int seed_kind, int init_nest, double amp_factor, lpm_mesh** out)
{
    if (!out || init_nest < 0 || init_nest > 12) return LPM_ERR_INVALID;
    const Seed* s = nullptr;
    switch (seed_kind) {
        case LPM_TRI_HEX_SEED: s = &kSeedHex; break;
        case LPM_QUAD_RECT_SEED: s = &kSeedQuad; break;
        case LPM_ICOS_TRI_SPHERE_SEED: s = &kSeedIcos; break;
        case LPM_CUBED_SPHERE_SEED: s = &kSeedCube; break;
        case LPM_BETA_PLANE_SEED: s = &kSeedBeta; break;
        default: return LPM_ERR_INVALID;   // the reference logs "invalid meshSeed" / "seed not implemented"
    }
    lpm_mesh* h = new lpm_mesh();
    h->m.initialize(*s, seed_kind, amp_factor);
    h->m.refine(init_nest);
    *out = h;
    return LPM_OK;
}

void lpm_mesh_destroy(lpm_mesh* h) { delete h; }

int64_t lpm_mesh_num_particles(const lpm_mesh* h) { return h->m.nParticles(); }
int64_t lpm_mesh_num_faces(const lpm_mesh* h) { return h->m.nFaces(); }
int64_t lpm_mesh_num_edges(const lpm_mesh* h) { return h->m.nEdges(); }

int64_t lpm_mesh_num_leaf_faces(const lpm_mesh* h)
{
    int64_t c = 0;
    for (uint8_t k : h->m.fHasKids) c += !k;
    return c;
}

int64_t lpm_mesh_num_leaf_edges(const lpm_mesh* h)
{
    int64_t c = 0;
    for (uint8_t k : h->m.eHasKids) c += !k;
    return c;
}

double lpm_mesh_max_edge_length(const lpm_mesh* h) { return h->m.maxEdgeLength(); }

int lpm_mesh_get_particles(const lpm_mesh* h, double* x, double* y, double* z, double* area, int32_t* is_active)
{
    const Mesh& m = h->m;
    size_t n = (size_t)m.nParticles();
    if (x) std::memcpy(x, m.x.data(), n * sizeof(double));
    if (y) std::memcpy(y, m.y.data(), n * sizeof(double));
    if (z) std::memcpy(z, m.z.data(), n * sizeof(double));
    if (area) std::memcpy(area, m.area.data(), n * sizeof(double));
    if (is_active) std::memcpy(is_active, m.active.data(), n * sizeof(int32_t));
    return LPM_OK;
}

int lpm_mesh_get_leaf_faces(const lpm_mesh* h, int32_t* verts, int32_t* center)
{
    const Mesh& m = h->m;
    int64_t c = 0;
    for (int64_t f = 0; f < m.nFaces(); ++f) {
        if (m.fHasKids[f]) continue;
        for (int j = 0; j < m.vpf; ++j) verts[c * m.vpf + j] = m.fVerts[f * m.vpf + j];
        center[c] = m.fCenter[f];
        ++c;
    }
    return LPM_OK;
}

// Legacy-format ASCII .vtk PolyData output of a mesh and point fields, the layout of the
// reference's outputVTKPrivate (src/SphereBVE.f90:283-328): header (src/OutputWriter.f90:272-283),
// POINTS (src/Particles.f90:368-390), POLYGONS -- every leaf face as vertsPerFace triangles
// (vertex j, vertex j+1, centre particle), 0-based (src/Faces.f90:388-408), POINT_DATA with
// lagParam (src/Particles.f90:419-436; the uniform meshes here have x0 == x at creation, so the
// mesh's own coordinates are the Lagrangian parameter) and the caller's fields
// (src/Field.f90:285-311: "SCALARS name  double nDim", scalars below ZERO_TOL = 1e-14 written as 0),
// CELL_DATA faceArea (src/Faces.f90:418-440).  Positions x, y, z may be NULL (the mesh's own);
// field f has ndim[f] components stored one after the other in data[f] (ndim[f] * N doubles).
// Numbers are written with 17 significant digits (the reference's list-directed output is
// compiler dependent; any VTK reader accepts both).
int lpm_mesh_write_vtk(const lpm_mesh* h, const char* filename, const char* title, const double* x, const double* y,
                       const double* z, int nfields, const char* const* names, const int* ndim,
                       const double* const* data)
{
    if (!h || !filename || nfields < 0 || (nfields > 0 && (!names || !ndim || !data))) return LPM_ERR_INVALID;
    const Mesh& m = h->m;
    const int64_t n = m.nParticles();
    for (int f = 0; f < nfields; ++f)
        if (!names[f] || !data[f] || ndim[f] < 1 || ndim[f] > 3) return LPM_ERR_INVALID;
    FILE* fp = std::fopen(filename, "w");
    if (!fp) return LPM_ERR_INVALID;
    const double* px = x ? x : m.x.data();
    const double* py = y ? y : m.y.data();
    const double* pz = z ? z : m.z.data();
    const bool planar = !m.sphere;
    std::fprintf(fp, "# vtk DataFile Version 2.0\n%s\nASCII\nDATASET POLYDATA\n", (title && *title) ? title : " ");
    std::fprintf(fp, "POINTS %8lld double \n", (long long)n);
    for (int64_t j = 0; j < n; ++j) std::fprintf(fp, "%.17g %.17g %.17g\n", px[j], py[j], planar ? 0.0 : pz[j]);
    int64_t nleaf = 0;
    for (uint8_t k : m.fHasKids) nleaf += !k;
    const int64_t ncells = (int64_t)m.vpf * nleaf;
    std::fprintf(fp, "POLYGONS %8lld   %8lld\n", (long long)ncells, (long long)(4 * ncells));
    for (int64_t f = 0; f < m.nFaces(); ++f) {
        if (m.fHasKids[f]) continue;
        for (int j = 0; j < m.vpf; ++j)
            std::fprintf(fp, "%10d%10d%10d%10d\n", 3, m.fVerts[f * m.vpf + j], m.fVerts[f * m.vpf + (j + 1) % m.vpf],
                         m.fCenter[f]);
    }
    std::fprintf(fp, "POINT_DATA %8lld\n", (long long)n);
    std::fprintf(fp, "SCALARS lagParam double 3\nLOOKUP_TABLE default\n");
    for (int64_t j = 0; j < n; ++j) std::fprintf(fp, "%.17g %.17g %.17g\n", m.x[j], m.y[j], planar ? 0.0 : m.z[j]);
    for (int f = 0; f < nfields; ++f) {
        std::fprintf(fp, "SCALARS %s  double %4d\nLOOKUP_TABLE default\n", names[f], ndim[f]);
        const double* d = data[f];
        for (int64_t j = 0; j < n; ++j) {
            if (ndim[f] == 1) {
                std::fprintf(fp, "%.17g\n", std::fabs(d[j]) < 1.0e-14 ? 0.0 : d[j]);
            } else if (ndim[f] == 2) {
                std::fprintf(fp, "%.17g %.17g\n", d[j], d[n + j]);
            } else {
                std::fprintf(fp, "%.17g %.17g %.17g\n", d[j], d[n + j], d[2 * n + j]);
            }
        }
    }
    std::fprintf(fp, "CELL_DATA %8lld\nSCALARS faceArea double 1\nLOOKUP_TABLE default\n", (long long)ncells);
    for (int64_t f = 0; f < m.nFaces(); ++f) {
        if (m.fHasKids[f]) continue;
        for (int j = 0; j < m.vpf; ++j) std::fprintf(fp, "%.17g\n", m.area[m.fCenter[f]]);
    }
    const bool ok = std::ferror(fp) == 0;
    return (std::fclose(fp) == 0 && ok) ? LPM_OK : LPM_ERR_INVALID;
}

}  // extern "C"
