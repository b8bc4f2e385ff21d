// MeshUDF marching cubes as plain C++17: the reference's only native component,
// src/marching_cubes/_marching_cubes_lewiner_cy.pyx (marching_cubes_udf :1116-1774, Cell :87-851, the_big_switch :1848,
// check_the_big_switch :2125, test_face :2404, test_internal :2436, compute_edge_vote :1777-1807), restated without Cython /
// numpy.  Host code by design: the traversal is a raster scan that seeds a three-queue breadth-first search whose visit ORDER
// decides the pseudo-signs (every vote reads signs written by earlier cells), so the result is only reproducible serially.
//
// What is kept bit for bit (tests/test_meshudf_port.py compares vertex / face / normal / value arrays with the reference build):
//   * the visit order, the vote arithmetic in float (no contraction: compile with -ffp-contract=off), the double arithmetic of
//     the vertex interpolation, the thresholds, the `>= 2 reused vertices` admission test of non-seed cells;
//   * Lewiner's case resolution (face tests, interior tests) and his look-up tables, loaded as DATA (diffudf_b200/data/
//     lewiner_luts.bin, written by tools/export_lewiner_luts.py);
//   * the reference's observable quirks: the centre vertex accumulates (sum_z, sum_y, 0) as its gradient (:843-848 assign
//     v12_xg twice and v12_zg never), an interior test that falls through returns 0 (:2553-2560), the anchor vector survives
//     from one cell to the next when all eight gradients vanish (:1381), neighbours are queued up to N - 3 only (:1412-1423).
// What is different: the 4 N^3-int vertex-reuse array (2.1 GB at 512^3, :195) is a hash map keyed by the same slot number,
// the case resolution is ONE function that returns (table, sub-index, triangle count) and serves both the counting pass and
// the emitting pass, the interior test reads its edge geometry from a table instead of twelve written-out branches.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <string>
#include <unordered_map>
#include <vector>

extern "C" {
struct dudf_meshudf_result {
  float* vertices;   // [n_vertices][3]  (x, y, z) in grid units, as Cell.get_vertices
  float* normals;    // [n_vertices][3]  normalised accumulated gradients, as Cell.get_normals
  float* values;     // [n_vertices]     as Cell.get_values
  int32_t* faces;    // [n_faces][3]     vertex indices in emission order, as Cell.get_faces reshaped (-1, 3)
  int64_t n_vertices, n_faces;
};
int dudf_meshudf_mc(const float* im, const float* grads, int nz, int ny, int nx, int step, float avg_thresh, float max_thresh,
                    const unsigned char* mask, const void* luts_blob, int64_t luts_bytes, dudf_meshudf_result* out);
void dudf_meshudf_free(dudf_meshudf_result* r);
}
void dudf_set_error(const char* fmt, ...);      // the library's error channel (dudf_api.cu): dudf_last_error()

namespace {

const double EPS = 2.220446049250313e-16;      // np.spacing(1.0), the reference's FLT_EPSILON (:37)

struct Lut {
  const int8_t* v = nullptr;
  int l1 = 1, l2 = 1;
  int at(int i) const { return v[i]; }
  int at(int i, int j) const { return v[i * l1 + j]; }
  int at(int i, int j, int k) const { return v[(i * l1 + j) * l2 + k]; }
};

struct Luts {
  std::unordered_map<std::string, Lut> t;
  const Lut& operator[](const char* name) const { return t.at(name); }
  bool parse(const unsigned char* p, int64_t n) {
    if (n < 12 || memcmp(p, "DUDFLUT1", 8) != 0) return false;
    uint32_t cnt;
    memcpy(&cnt, p + 8, 4);
    int64_t off = 12;
    for (uint32_t i = 0; i < cnt; ++i) {
      if (off + 36 > n) return false;
      char name[17] = {0};
      memcpy(name, p + off, 16);
      uint32_t h[5];
      memcpy(h, p + off + 16, 20);
      off += 36;
      if (off + h[4] > n) return false;
      Lut l;
      l.v = reinterpret_cast<const int8_t*>(p + off);
      l.l1 = h[0] > 1 ? (int)h[2] : 1;
      l.l2 = h[0] > 2 ? (int)h[3] : 1;
      t[name] = l;
      off += (h[4] + 3) & ~3u;
    }
    static const char* need[] = {"EDGESRELX", "EDGESRELY", "EDGESRELZ", "CASES", "TILING1", "TILING14", "TEST13", "SUBCONFIG13"};
    for (const char* nm : need)
      if (!t.count(nm)) return false;
    return true;
  }
};

// which triangles a cell gets: rows `nt` triangles of table `lut` at [config] or [config][sub]
struct Tiling {
  const Lut* lut = nullptr;
  int sub = -1;
  int nt = 0;
};

struct Cube {
  const Luts& L;
  int nx, ny, nz;
  int x = 0, y = 0, z = 0, step = 1;
  double v[8] = {0};            // corner values v0..v7 in the paper's numbering
  double vv[8] = {0};           // the same, indexed by dz*4 + dy*2 + dx
  double vg[24] = {0};          // corner gradients (finite differences along the cube edges)
  double vmax = 0;
  double c12[3] = {0}, c12g[3] = {0};   // centre vertex and its gradient
  bool c12_done = false;
  int index = 0;
  std::unordered_map<int64_t, int> slot;      // vertex-reuse table: 4 slots per cell (edge 0, edge 3, edge 8, centre)
  std::vector<float> verts, norms, vals;
  std::vector<int32_t> faces;

  Cube(const Luts& l, int nx_, int ny_, int nz_) : L(l), nx(nx_), ny(ny_), nz(nz_) {}

  void set(int x_, int y_, int z_, int st, const double* c) {
    x = x_; y = y_; z = z_; step = st;
    index = 0;
    for (int i = 0; i < 8; ++i) {
      v[i] = c[i];
      if (v[i] > 0.0) index |= 1 << i;
    }
    c12_done = false;
  }

  int64_t slot_of(int e) const {
    int64_t i = (int64_t)ny * nx * z + (int64_t)nx * y + x;
    int j = 0, up = 0;
    if (e < 8) {                       // horizontal edges: 0-3 in this layer, 4-7 are edges 0-3 of the layer above
      if (e >= 4) { e -= 4; up = 1; }
      if (e == 1) { i += step; j = 1; }
      else if (e == 2) i += (int64_t)nx * step;
      else if (e == 3) j = 1;
    } else if (e < 12) {               // vertical edges
      j = 2;
      if (e == 9) i += step;
      else if (e == 10) i += (int64_t)nx * step + step;
      else if (e == 11) i += (int64_t)nx * step;
    } else {
      j = 3;                           // centre vertex
    }
    i += (int64_t)nx * ny * up;
    return 4 * i + j;
  }
  int lookup(int64_t s) const {
    auto it = slot.find(s);
    return it == slot.end() ? -1 : it->second;
  }

  void prepare() {
    static const int perm[8] = {0, 1, 3, 2, 4, 5, 7, 6};
    for (int i = 0; i < 8; ++i) vv[i] = v[perm[i]];
    double lo = 0.0, hi = 0.0;
    for (int i = 0; i < 8; ++i) {
      if (vv[i] > hi) hi = vv[i];
      if (vv[i] < lo) lo = vv[i];
    }
    vmax = hi - lo;
    // per corner: the x, y, z differences along the three cube edges that meet in it, always low corner minus high corner
    static const int gx[8][2] = {{0, 1}, {0, 1}, {3, 2}, {3, 2}, {4, 5}, {4, 5}, {7, 6}, {7, 6}};
    static const int gy[8][2] = {{0, 3}, {1, 2}, {1, 2}, {0, 3}, {4, 7}, {5, 6}, {5, 6}, {4, 7}};
    static const int gz[8][2] = {{0, 4}, {1, 5}, {2, 6}, {3, 7}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
    for (int i = 0; i < 8; ++i) {
      vg[i * 3 + 0] = v[gx[i][0]] - v[gx[i][1]];
      vg[i * 3 + 1] = v[gy[i][0]] - v[gy[i][1]];
      vg[i * 3 + 2] = v[gz[i][0]] - v[gz[i][1]];
    }
  }

  void centre() {
    static const double cx[8] = {0, 1, 1, 0, 0, 1, 1, 0}, cy[8] = {0, 0, 1, 1, 0, 0, 1, 1}, cz[8] = {0, 0, 0, 0, 1, 1, 1, 1};
    double w[8], fx = 0.0, fy = 0.0, fz = 0.0, ff = 0.0;
    for (int i = 0; i < 8; ++i) w[i] = 1.0 / (EPS + std::fabs(v[i]));
    for (int i = 0; i < 8; ++i) { fx += cx[i] * w[i]; fy += cy[i] * w[i]; fz += cz[i] * w[i]; ff += w[i]; }
    const double stp = (double)step;
    c12[0] = x + stp * fx / ff;
    c12[1] = y + stp * fy / ff;
    c12[2] = z + stp * fz / ff;
    double s[3];
    for (int k = 0; k < 3; ++k)
      s[k] = w[0] * vg[0 * 3 + k] + w[1] * vg[1 * 3 + k] + w[2] * vg[2 * 3 + k] + w[3] * vg[3 * 3 + k] + w[4] * vg[4 * 3 + k] +
             w[5] * vg[5 * 3 + k] + w[6] * vg[6 * 3 + k] + w[7] * vg[7 * 3 + k];
    c12g[0] = s[2];       // the reference stores the z sum in the x slot, keeps the y sum and never sets the z slot
    c12g[1] = s[1];
    c12g[2] = 0.0;
    c12_done = true;
  }

  int new_vertex(float px, float py, float pz) {
    verts.push_back(px); verts.push_back(py); verts.push_back(pz);
    norms.push_back(0.f); norms.push_back(0.f); norms.push_back(0.f);
    vals.push_back(0.f);
    return (int)vals.size() - 1;
  }
  void add_grad(int vi, float gx, float gy, float gz) {
    norms[vi * 3 + 0] += gx; norms[vi * 3 + 1] += gy; norms[vi * 3 + 2] += gz;
  }
  void add_corner_grad(int vi, int corner, float strength) {
    add_grad(vi, (float)(vg[corner * 3 + 0] * strength), (float)(vg[corner * 3 + 1] * strength), (float)(vg[corner * 3 + 2] * strength));
  }
  void add_face(int vi) {
    faces.push_back(vi);
    if (vmax > vals[vi]) vals[vi] = (float)vmax;
  }

  // one triangle corner on edge e (12 = centre): reuse the vertex of that edge if it exists, else interpolate a new one
  void emit(int e) {
    const int64_t s = slot_of(e);
    int vi = lookup(s);
    if (e == 12) {
      if (!c12_done) centre();
      if (vi < 0) {
        vi = new_vertex((float)c12[0], (float)c12[1], (float)c12[2]);
        slot[s] = vi;
      }
      add_face(vi);
      add_grad(vi, (float)c12g[0], (float)c12g[1], (float)c12g[2]);
      return;
    }
    const int dx1 = L["EDGESRELX"].at(e, 0), dx2 = L["EDGESRELX"].at(e, 1);
    const int dy1 = L["EDGESRELY"].at(e, 0), dy2 = L["EDGESRELY"].at(e, 1);
    const int dz1 = L["EDGESRELZ"].at(e, 0), dz2 = L["EDGESRELZ"].at(e, 1);
    const int i1 = dz1 * 4 + dy1 * 2 + dx1, i2 = dz2 * 4 + dy2 * 2 + dx2;
    const double w1 = 1.0 / (EPS + std::fabs(vv[i1])), w2 = 1.0 / (EPS + std::fabs(vv[i2]));
    if (vi < 0) {
      double fx = 0.0, fy = 0.0, fz = 0.0, ff = 0.0;
      fx += (double)dx1 * w1; fy += (double)dy1 * w1; fz += (double)dz1 * w1; ff += w1;
      fx += (double)dx2 * w2; fy += (double)dy2 * w2; fz += (double)dz2 * w2; ff += w2;
      const double stp = (double)step;
      vi = new_vertex((float)((double)x + stp * fx / ff), (float)((double)y + stp * fy / ff), (float)((double)z + stp * fz / ff));
      slot[s] = vi;
    }
    add_face(vi);
    add_corner_grad(vi, i1, (float)w1);
    add_corner_grad(vi, i2, (float)w2);
  }

  int edge_of(const Tiling& t, int config, int k) const { return t.sub < 0 ? t.lut->at(config, k) : t.lut->at(config, t.sub, k); }

  void emit_all(const Tiling& t, int config) {
    if (!t.lut) return;
    prepare();
    for (int k = 0; k < 3 * t.nt; ++k) emit(edge_of(t, config, k));
  }
  // number of DISTINCT already existing vertices among the triangle corners of the tiling
  int count_existing(const Tiling& t, int config) {
    if (!t.lut) return 0;
    prepare();
    std::vector<int> seen;
    int n = 0;
    for (int k = 0; k < 3 * t.nt; ++k) {
      const int vi = lookup(slot_of(edge_of(t, config, k)));
      bool dup = false;
      for (int s : seen) dup = dup || (s == vi);
      if (!dup && vi >= 0) ++n;
      seen.push_back(vi);
    }
    return n;
  }

  // ---- Lewiner's ambiguity tests ----
  bool test_face(int face) const {
    static const int corners[7][4] = {{0, 0, 0, 0}, {0, 4, 5, 1}, {1, 5, 6, 2}, {2, 6, 7, 3}, {3, 7, 4, 0}, {0, 3, 2, 1}, {4, 7, 6, 5}};
    const int af = face < 0 ? -face : face;
    double A = 0, B = 0, C = 0, D = 0;
    if (af >= 1 && af <= 6) { A = v[corners[af][0]]; B = v[corners[af][1]]; C = v[corners[af][2]]; D = v[corners[af][3]]; }
    const double acbd = A * C - B * D;
    if (acbd > -EPS && acbd < EPS) return face >= 0;
    return face * A * acbd >= 0;
  }
  int test_interior(int kase, int config, int subconfig, int s) const {
    double At = 0.0, Bt = 0.0, Ct = 0.0, Dt = 0.0;
    if (kase == 4 || kase == 10) {
      const double a = (v[4] - v[0]) * (v[6] - v[2]) - (v[7] - v[3]) * (v[5] - v[1]);
      const double b = v[2] * (v[4] - v[0]) + v[0] * (v[6] - v[2]) - v[1] * (v[7] - v[3]) - v[3] * (v[5] - v[1]);
      const double t = -b / (2 * a + EPS);
      if (t < 0 || t > 1) return s > 0;
      At = v[0] + (v[4] - v[0]) * t;
      Bt = v[3] + (v[7] - v[3]) * t;
      Ct = v[2] + (v[6] - v[2]) * t;
      Dt = v[1] + (v[5] - v[1]) * t;
    } else if (kase == 6 || kase == 7 || kase == 12 || kase == 13) {
      int edge = -1;
      if (kase == 6) edge = L["TEST6"].at(config, 2);
      else if (kase == 7) edge = L["TEST7"].at(config, 4);
      else if (kase == 12) edge = L["TEST12"].at(config, 3);
      else edge = L["TILING13_5_1"].at(config, subconfig, 0);
      // reference edge p -> q and the three edges parallel to it, walking round the cube (B, C, D start corners and ends)
      static const int geo[12][8] = {{0, 1, 3, 2, 7, 6, 4, 5}, {1, 2, 0, 3, 4, 7, 5, 6}, {2, 3, 1, 0, 5, 4, 6, 7}, {3, 0, 2, 1, 6, 5, 7, 4},
                                     {4, 5, 7, 6, 3, 2, 0, 1}, {5, 6, 4, 7, 0, 3, 1, 2}, {6, 7, 5, 4, 1, 0, 2, 3}, {7, 4, 6, 5, 2, 1, 3, 0},
                                     {0, 4, 3, 7, 2, 6, 1, 5}, {1, 5, 0, 4, 3, 7, 2, 6}, {2, 6, 1, 5, 0, 4, 3, 7}, {3, 7, 2, 6, 1, 5, 0, 4}};
      if (edge >= 0 && edge < 12) {
        const int* g = geo[edge];
        const double t = v[g[0]] / (v[g[0]] - v[g[1]] + EPS);
        At = 0;
        Bt = v[g[2]] + (v[g[3]] - v[g[2]]) * t;
        Ct = v[g[4]] + (v[g[5]] - v[g[4]]) * t;
        Dt = v[g[6]] + (v[g[7]] - v[g[6]]) * t;
      }
    }
    int test = 0;
    if (At >= 0) test += 1;
    if (Bt >= 0) test += 2;
    if (Ct >= 0) test += 4;
    if (Dt >= 0) test += 8;
    switch (test) {
      case 5: return (At * Ct - Bt * Dt < EPS) ? (s > 0) : 0;
      case 10: return (At * Ct - Bt * Dt >= EPS) ? (s > 0) : 0;
      case 7: case 11: case 13: case 14: case 15: return s < 0;
      default: return s > 0;
    }
  }

  // Lewiner's case table walk: which tiling the (case, config) of the current cube resolves to
  Tiling resolve(int kase, int config) const {
    auto T = [&](const char* name, int nt, int sub = -1) { Tiling t; t.lut = &L[name]; t.sub = sub; t.nt = nt; return t; };
    switch (kase) {
      case 1: return T("TILING1", 1);
      case 2: return T("TILING2", 2);
      case 3: return test_face(L["TEST3"].at(config)) ? T("TILING3_2", 4) : T("TILING3_1", 2);
      case 4: return test_interior(4, config, 0, L["TEST4"].at(config)) ? T("TILING4_1", 2) : T("TILING4_2", 6);
      case 5: return T("TILING5", 3);
      case 6:
        if (test_face(L["TEST6"].at(config, 0))) return T("TILING6_2", 5);
        return test_interior(6, config, 0, L["TEST6"].at(config, 1)) ? T("TILING6_1_1", 3) : T("TILING6_1_2", 9);
      case 7: {
        int sub = 0;
        for (int k = 0; k < 3; ++k)
          if (test_face(L["TEST7"].at(config, k))) sub += 1 << k;
        switch (sub) {
          case 0: return T("TILING7_1", 3);
          case 1: return T("TILING7_2", 5, 0);
          case 2: return T("TILING7_2", 5, 1);
          case 4: return T("TILING7_2", 5, 2);
          case 3: return T("TILING7_3", 9, 0);
          case 5: return T("TILING7_3", 9, 1);
          case 6: return T("TILING7_3", 9, 2);
          default: return test_interior(7, config, 7, L["TEST7"].at(config, 3)) ? T("TILING7_4_2", 9) : T("TILING7_4_1", 5);
        }
      }
      case 8: return T("TILING8", 2);
      case 9: return T("TILING9", 4);
      case 10:
      case 12: {
        const char* test = kase == 10 ? "TEST10" : "TEST12";
        const bool f0 = test_face(L[test].at(config, 0));
        const bool f1 = test_face(L[test].at(config, 1));
        if (f0 && f1) return T(kase == 10 ? "TILING10_1_1_" : "TILING12_1_1_", 4);
        if (f0) return T(kase == 10 ? "TILING10_2" : "TILING12_2", 8);
        if (f1) return T(kase == 10 ? "TILING10_2_" : "TILING12_2_", 8);
        if (test_interior(kase, config, 0, L[test].at(config, 2))) return T(kase == 10 ? "TILING10_1_1" : "TILING12_1_1", 4);
        return T(kase == 10 ? "TILING10_1_2" : "TILING12_1_2", 8);
      }
      case 11: return T("TILING11", 4);
      case 13: {
        int bits = 0;
        for (int k = 0; k < 6; ++k)
          if (test_face(L["TEST13"].at(config, k))) bits += 1 << k;
        const int sub = L["SUBCONFIG13"].at(bits);
        if (sub == 0) return T("TILING13_1", 4);
        if (sub >= 1 && sub <= 6) return T("TILING13_2", 6, sub - 1);
        if (sub >= 7 && sub <= 18) return T("TILING13_3", 10, sub - 7);
        if (sub >= 19 && sub <= 22) return T("TILING13_4", 12, sub - 19);
        if (sub >= 23 && sub <= 26)
          return test_interior(13, config, sub - 23, L["TEST13"].at(config, 6)) ? T("TILING13_5_1", 6, sub - 23) : T("TILING13_5_2", 10, sub - 23);
        if (sub >= 27 && sub <= 38) return T("TILING13_3_", 10, sub - 27);
        if (sub >= 39 && sub <= 44) return T("TILING13_2_", 6, sub - 39);
        if (sub == 45) return T("TILING13_1_", 4);
        return Tiling();
      }
      case 14: return T("TILING14", 4);
    }
    return Tiling();
  }
};

inline float sgn(float a) { return a > 0 ? 1.f : (a < 0 ? -1.f : 0.f); }
inline float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
inline bool nonzero(const float* a) { return (std::fabs(a[0]) + std::fabs(a[1]) + std::fabs(a[2])) > 0; }

// vote of a signed neighbour along one grid axis (dz, dy, dx: exactly one non-zero).  grads are stored (z, y, x)-major with the
// component order of the caller; the component tested is the one that belongs to the axis, as in the reference
inline float edge_vote(const float* g1, const float* g2, float dz, float dy, float dx) {
  const float dsum = dz + dy + dx;
  float p1, p2;
  if (dz != 0) { p1 = g1[0]; p2 = g2[0]; }
  else if (dy != 0) { p1 = g1[1]; p2 = g2[1]; }
  else { p1 = g1[2]; p2 = g2[2]; }
  if (dsum > 0) return (p2 > 0 && p1 < 0) ? 1.0f : dot3(g1, g2);
  return (p2 < 0 && p1 > 0) ? 1.0f : dot3(g1, g2);
}

struct Tup { int z, y, x; };

int run(const float* im, const float* grads, int Nz, int Ny, int Nx, int st, float avg_thresh, float max_thresh, const unsigned char* mask,
        const Luts& luts, dudf_meshudf_result* out) {
  const int64_t sy = Nx, sz = (int64_t)Nx * Ny, total = sz * Nz;
  auto at = [&](int z, int y, int x) { return z * sz + y * sy + x; };
  const double voxel = 2.0 / (Nx - 1);
  const float avg_lim = (float)((double)avg_thresh * voxel), max_lim = (float)((double)max_thresh * voxel);
  const int bx = Nx - 2 * st, by = Ny - 2 * st, bz = Nz - 2 * st;
  const float unsure = 0.707f;
  std::vector<float> sign(total, 0.f);
  std::vector<unsigned char> known(total, 0), visited(total, 0);
  Cube cube(luts, Nx, Ny, Nz);
  std::deque<Tup> q, q_unsure, q_complex;
  float base[3] = {0.f, 0.f, 0.f};
  const int dirz[6] = {st, -st, 0, 0, 0, 0}, diry[6] = {0, 0, st, -st, 0, 0}, dirx[6] = {0, 0, 0, 0, st, -st};

  auto push_neighbours = [&](int z, int y, int x) {
    if (x + st < bx) q.push_back({z, y, x + st});
    if (y + st < by) q.push_back({z, y + st, x});
    if (x - st >= 0) q.push_back({z, y, x - st});
    if (y - st >= 0) q.push_back({z, y - st, x});
    if (z - st >= 0) q.push_back({z - st, y, x});
    if (z + st < bz) q.push_back({z + st, y, x});
  };
  auto thin_enough = [&](const int64_t* c) {
    const float s = im[c[0]] + im[c[1]] + im[c[2]] + im[c[3]] + im[c[4]] + im[c[5]] + im[c[6]] + im[c[7]];
    const float avg = (float)(0.125 * s);
    float mx = im[c[7]];
    for (int i = 6; i >= 0; --i) mx = std::max(im[c[i]], mx);
    return avg < avg_lim && mx <= max_lim;
  };

  // One cell.  seed: reached by the raster scan (always emits).  Otherwise a cell of the breadth-first search:
  // `full` = false while the neighbours of an unsure cell are revisited (signs only, no faces).  Returns true when the seed
  // emitted faces (the search starts from it).
  auto cell = [&](int z, int y, int x, bool seed, bool full) -> bool {
    const int zs = z + st, ys = y + st, xs = x + st;
    if (visited[at(z, y, x)] || !(mask == nullptr || mask[at(zs, ys, xs)])) return false;
    const int cz[8] = {z, z, z, z, zs, zs, zs, zs}, cy[8] = {y, y, ys, ys, y, y, ys, ys}, cx[8] = {x, xs, xs, x, x, xs, xs, x};
    int64_t c[8];
    for (int i = 0; i < 8; ++i) c[i] = at(cz[i], cy[i], cx[i]);
    if (!thin_enough(c)) return false;
    float votes[8];
    int nvotes[8];
    for (int vi = 0; vi < 8; ++vi) {
      nvotes[vi] = 0;
      votes[vi] = 0.f;
      if (known[c[vi]]) { nvotes[vi] = 1; votes[vi] = sign[c[vi]]; continue; }
      if (im[c[vi]] == 0.0f) { nvotes[vi] = 1; continue; }
      for (int d = 0; d < 6; ++d) {
        int i = 0, reach = 1;
        while (i < reach) {
          ++i;
          const int z2 = cz[vi] + i * dirz[d], y2 = cy[vi] + i * diry[d], x2 = cx[vi] + i * dirx[d];
          if (z2 > bz || z2 < 0 || y2 > by || y2 < 0 || x2 > bx || x2 < 0) break;
          const int64_t o = at(z2, y2, x2);
          if (im[o] == 0.0f) {               // an exact zero does not vote: look one vertex further
            if (i >= reach) ++reach;
            continue;
          }
          if (sign[o] == 0.0f) continue;      // not signed yet
          nvotes[vi] += 1;
          votes[vi] += sign[o] * edge_vote(grads + 3 * c[vi], grads + 3 * o, (float)dirz[d], (float)diry[d], (float)dirx[d]);
        }
      }
      if (!seed && nvotes[vi] >= 1 && std::fabs(votes[vi]) / nvotes[vi] < unsure && !q.empty()) {
        if (full) q_unsure.push_back({z, y, x});
        return false;
      }
      sign[c[vi]] = sgn(votes[vi]);         // provisional: used by the neighbours, recomputed until the cell is `known`
    }
    bool all_voted = true;
    for (int vi = 0; vi < 8; ++vi) all_voted = all_voted && nvotes[vi] >= 1;
    if (!all_voted) {
      // anchor direction: the first signed corner with a gradient, else the first corner with a gradient (order 0 1 3 2 4 5 7 6)
      static const int order[8] = {0, 1, 3, 2, 4, 5, 7, 6};
      double anchor = 1.0;
      int pick = -1;
      for (int k = 0; k < 8 && pick < 0; ++k)
        if (known[c[order[k]]] && nonzero(grads + 3 * c[order[k]])) { pick = order[k]; anchor = sgn(sign[c[pick]]); }
      for (int k = 0; k < 8 && pick < 0; ++k)
        if (nonzero(grads + 3 * c[order[k]])) pick = order[k];
      if (pick >= 0)
        for (int k = 0; k < 3; ++k) base[k] = grads[3 * c[pick] + k];
      for (int k = 0; k < 3; ++k) base[k] = (float)(anchor * base[k]);
      const bool careful = !seed && full && !q.empty();
      for (int vi = 0; vi < 8; ++vi) {
        if (nvotes[vi] != 0) continue;
        const float dv = dot3(base, grads + 3 * c[vi]);
        if (careful) {
          votes[vi] = dv;
          if (std::fabs(dv) < unsure) { q_unsure.push_back({z, y, x}); return false; }
        }
        sign[c[vi]] = sgn(dv);
      }
    }
    if (!seed && !full) return false;
    double val[8];
    for (int i = 0; i < 8; ++i) val[i] = sign[c[i]] * im[c[i]];
    cube.set(x, y, z, st, val);
    for (int i = 0; i < 8; ++i) known[c[i]] = 1;
    const int kase = luts["CASES"].at(cube.index, 0);
    if (kase <= 0) { visited[at(z, y, x)] = 1; return false; }
    if (!seed) {
      const bool simple = kase == 1 || kase == 2 || kase == 5 || kase == 8 || kase == 9;
      if (!simple && (!q.empty() || !q_unsure.empty())) { q_complex.push_back({z, y, x}); return false; }
    }
    const int config = luts["CASES"].at(cube.index, 1);
    const Tiling t = cube.resolve(kase, config);
    if (!seed && cube.count_existing(t, config) < 2) return false;
    visited[at(z, y, x)] = 1;
    cube.emit_all(t, config);
    push_neighbours(z, y, x);
    return true;
  };

  // raster scan exactly as the reference's `zi = -st; while zi < bound: zi += st` loops run it
  for (int z = 0; z - st < bz; z += st) {
    for (int y = 0; y - st < by; y += st) {
      for (int x = 0; x - st < bx; x += st) {
        if (!cell(z, y, x, true, true)) continue;
        bool full = true;
        while (!q.empty() || !q_unsure.empty() || !q_complex.empty()) {
          Tup cur;
          if (q.empty()) {
            if (q_unsure.empty()) {
              cur = q_complex.front();
              q_complex.pop_front();
            } else {
              cur = q_unsure.front();
              if (full) {       // first the neighbours of the unsure cell (signs only), then the cell itself
                if (visited[at(cur.z, cur.y, cur.x)]) { q_unsure.pop_front(); continue; }
                push_neighbours(cur.z, cur.y, cur.x);
                full = false;
                continue;
              }
              q_unsure.pop_front();
              full = true;
            }
          } else {
            cur = q.front();
            q.pop_front();
          }
          cell(cur.z, cur.y, cur.x, false, full);
        }
      }
    }
  }

  const int64_t nv = (int64_t)cube.vals.size(), nf = (int64_t)cube.faces.size() / 3;
  out->n_vertices = nv;
  out->n_faces = nf;
  out->vertices = (float*)malloc(sizeof(float) * 3 * (nv ? nv : 1));
  out->normals = (float*)malloc(sizeof(float) * 3 * (nv ? nv : 1));
  out->values = (float*)malloc(sizeof(float) * (nv ? nv : 1));
  out->faces = (int32_t*)malloc(sizeof(int32_t) * 3 * (nf ? nf : 1));
  if (!out->vertices || !out->normals || !out->values || !out->faces) { dudf_set_error("dudf_meshudf_mc: out of memory"); return 1; }
  memcpy(out->vertices, cube.verts.data(), sizeof(float) * 3 * nv);
  memcpy(out->values, cube.vals.data(), sizeof(float) * nv);
  memcpy(out->faces, cube.faces.data(), sizeof(int32_t) * 3 * nf);
  for (int64_t i = 0; i < nv; ++i) {
    double len = 0.0;
    for (int k = 0; k < 3; ++k) { const double d = cube.norms[i * 3 + k]; len += d * d; }
    if (len > 0.0) len = 1.0 / std::pow(len, 0.5);
    for (int k = 0; k < 3; ++k) out->normals[i * 3 + k] = (float)(cube.norms[i * 3 + k] * len);
  }
  return 0;
}

}  // namespace

extern "C" int dudf_meshudf_mc(const float* im, const float* grads, int nz, int ny, int nx, int step, float avg_thresh, float max_thresh,
                               const unsigned char* mask, const void* luts_blob, int64_t luts_bytes, dudf_meshudf_result* out) {
  if (!im || !grads || !luts_blob || !out) { dudf_set_error("dudf_meshudf_mc: null argument"); return 2; }
  if (nz < 2 || ny < 2 || nx < 2) { dudf_set_error("dudf_meshudf_mc: the volume must be at least 2 x 2 x 2"); return 2; }
  if (step < 1) { dudf_set_error("dudf_meshudf_mc: step must be at least one"); return 2; }
  memset(out, 0, sizeof(*out));
  Luts luts;
  if (!luts.parse((const unsigned char*)luts_blob, luts_bytes)) { dudf_set_error("dudf_meshudf_mc: malformed look-up table blob"); return 2; }
  try {
    return run(im, grads, nz, ny, nx, step, avg_thresh, max_thresh, mask, luts, out);
  } catch (const std::exception& e) {
    dudf_set_error("dudf_meshudf_mc: %s", e.what());
    return 1;
  }
}

extern "C" void dudf_meshudf_free(dudf_meshudf_result* r) {
  if (!r) return;
  free(r->vertices); free(r->normals); free(r->values); free(r->faces);
  memset(r, 0, sizeof(*r));
}
