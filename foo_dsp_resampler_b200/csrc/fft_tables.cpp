// fft_tables.cpp -- see fft_tables.hpp.
#include "fft_tables.hpp"
#include "pk_plan.hpp"

#include <algorithm>
#include <cassert>
#include <cmath>

namespace b200rate {

namespace {
constexpr double kPi = 3.14159265358979323846;
constexpr double kSqrtHalf = 0.70710678118654752440;

struct TreeWalk {
  CfftHostSched *s;
  std::vector<std::vector<uint16_t>> per_level;
  std::vector<std::vector<uint16_t>> quarter_children;
  void visit(int size, int off, bool quarter = false)
  {
    if (size == 16) { s->leaf16_off.push_back(static_cast<uint16_t>(off)); return; }
    if (size == 8) { s->leaf8_off.push_back(static_cast<uint16_t>(off)); return; }
    // fft(S) = fft(S/2) on the first half, fft(S/4) on each remaining quarter, then one pass
    visit(size >> 1, off);
    visit(size >> 2, off + (size >> 1), true);
    visit(size >> 2, off + 3 * (size >> 2), true);
    int lg = 0;
    while ((1 << lg) < size) ++lg;
    per_level[lg].push_back(static_cast<uint16_t>(off));
    if (quarter) quarter_children[lg].push_back(static_cast<uint16_t>(off));
  }
};
}  // namespace

int split_radix_index(int i, int n, int inverse)
{
  if (n <= 2) return i & 1;
  const int half = n >> 1, quarter = n >> 2;
  if (!(i & half)) return 2 * split_radix_index(i, half, inverse);
  const int sub = split_radix_index(i, quarter, inverse);
  const int upper = (i & quarter) != 0;
  return 4 * sub + (inverse != upper ? 1 : -1);
}

CfftHostSched build_cfft_sched(int bits)
{
  assert(bits >= 5 && bits <= 16);
  CfftHostSched s;
  s.bits = bits;
  const int m = 1 << bits;
  TreeWalk w{&s, std::vector<std::vector<uint16_t>>(17), std::vector<std::vector<uint16_t>>(17)};
  w.visit(m, 0);
  for (int lg = 5; lg <= bits; ++lg) {
    s.level_begin[lg] = static_cast<int>(s.node_off.size());
    s.level_cnt[lg] = static_cast<int>(w.per_level[lg].size());
    s.node_off.insert(s.node_off.end(), w.per_level[lg].begin(), w.per_level[lg].end());
    s.pyr_off[lg] = s.pyr_len;
    s.pyr_len += (1 << (lg - 2)) + 1;
  }
  for (int lg = 5; lg <= bits; ++lg) {
    s.qchild_begin[lg] = static_cast<int>(s.node_off.size());
    s.qchild_cnt[lg] = static_cast<int>(w.quarter_children[lg].size());
    s.node_off.insert(s.node_off.end(), w.quarter_children[lg].begin(), w.quarter_children[lg].end());
  }
  const int n16 = static_cast<int>(s.leaf16_off.size()), n8 = static_cast<int>(s.leaf8_off.size());
  for (int inv = 0; inv < 2; ++inv) {
    std::vector<int> natural(static_cast<size_t>(m));   // permuted[p] = natural_order[natural[p]]
    for (int p = 0; p < m; ++p) natural[p] = (-split_radix_index(p, m, inv)) & (m - 1);
    s.gather16[inv].assign(static_cast<size_t>(16) * n16, 0);
    s.gather8[inv].assign(static_cast<size_t>(8) * n8, 0);
    for (int t = 0; t < n16; ++t)
      for (int e = 0; e < 16; ++e)
        s.gather16[inv][static_cast<size_t>(e) * n16 + t] = static_cast<uint16_t>(natural[s.leaf16_off[t] + e]);
    for (int t = 0; t < n8; ++t)
      for (int e = 0; e < 8; ++e)
        s.gather8[inv][static_cast<size_t>(e) * n8 + t] = static_cast<uint16_t>(natural[s.leaf8_off[t] + e]);
  }
  return s;
}

int pk_slot(int p) { return p + (p >> 4) + (p >> 8); }

namespace {
// Order `offs` so that every aligned run of eight entries falls into eight different 16-byte bank groups
// (as far as the multiset allows): repeatedly take one entry from each of the fullest groups.
std::vector<uint16_t> spread_over_bank_groups(const std::vector<uint16_t> &offs)
{
  std::vector<std::vector<uint16_t>> bucket(8);
  for (uint16_t o : offs) bucket[static_cast<size_t>(pk_slot(o) & 7)].push_back(o);
  std::vector<uint16_t> out;
  while (out.size() < offs.size()) {
    int order[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    for (int a = 0; a < 8; ++a)                    // fullest buckets first (stable selection sort, 8 entries)
      for (int b = a + 1; b < 8; ++b)
        if (bucket[static_cast<size_t>(order[b])].size() > bucket[static_cast<size_t>(order[a])].size()) std::swap(order[a], order[b]);
    for (int a = 0; a < 8; ++a) {
      std::vector<uint16_t> &bk = bucket[static_cast<size_t>(order[a])];
      if (bk.empty()) continue;
      out.push_back(bk.back());
      bk.pop_back();
    }
  }
  return out;
}

// As above, but never lets two entries of an aligned run of eight share a bank group: entry k of every group goes
// to run k, the runs are padded with holes (kPkHole) -- idle lanes cost nothing, a two-way conflict costs a
// wavefront for each of the 16 (8) accesses of a leaf. The split-radix tree puts the 16-point leaves of a quarter
// at offsets 16 m with m in {0, 2, 3, 4, 6} (mod 8) plus one 7, i.e. only six distinct groups: without holes every
// run of eight lanes collides (profiles/README.md, round 2).
std::vector<uint16_t> spread_with_holes(const std::vector<uint16_t> &offs)
{
  std::vector<std::vector<uint16_t>> bucket(8);
  for (uint16_t o : offs) bucket[static_cast<size_t>(pk_slot(o) & 7)].push_back(o);
  size_t runs = 0;
  for (const auto &b : bucket) runs = std::max(runs, b.size());
  std::vector<uint16_t> out;
  for (size_t r = 0; r < runs; ++r) {
    size_t placed = 0;
    for (const auto &b : bucket)
      if (r < b.size()) { out.push_back(b[r]); ++placed; }
    if (r + 1 < runs) for (; placed < 8; ++placed) out.push_back(kPkHole);
  }
  return out;
}
}  // namespace

PkHostSched build_pk_sched(const CfftHostSched &h)
{
  PkHostSched s;
  const int bits = h.bits, m = 1 << bits;
  const std::vector<uint16_t> l16 = spread_over_bank_groups(h.leaf16_off), l8 = spread_over_bank_groups(h.leaf8_off);
  assert(static_cast<int>(l16.size()) == pk_n16(bits) && static_cast<int>(l8.size()) == pk_n8(bits));
  s.tasks = l16;
  const int n8p = pk_n8p(bits);
  std::vector<uint16_t> pairs(static_cast<size_t>(2 * n8p), 0xffff);
  for (size_t k = 0; k < l8.size(); ++k) {         // task t: entries t and t + n8p of the spread list
    const size_t t = k < static_cast<size_t>(n8p) ? k : k - static_cast<size_t>(n8p);
    pairs[2 * t + (k < static_cast<size_t>(n8p) ? 0 : 1)] = l8[k];
  }
  s.tasks.insert(s.tasks.end(), pairs.begin(), pairs.end());
  const PkPhaseList pl = pk_phase_list(bits);
  for (int ph = 0; ph < pl.n; ++ph) {
    const int lg = pl.lg[ph], depth = pl.depth[ph], q = 1 << (lg - 2);
    assert(static_cast<int>(s.tasks.size()) == pk_phase_base(bits, ph));
    auto emit = [&](int begin, int count) {
      for (int node = 0; node < count; ++node)
        for (int k = 0; k < q; ++k) s.tasks.push_back(static_cast<uint16_t>(h.node_off[static_cast<size_t>(begin + node)] + k));
    };
    const int top = lg + depth - 1;                // main tasks: every node of the phase's top size
    emit(h.level_begin[top], h.level_cnt[top]);
    assert(static_cast<int>(s.tasks.size()) == pk_phase_base(bits, ph) + pk_phase_main(bits, lg, depth));
    if (depth > 1) emit(h.qchild_begin[top - 1], h.qchild_cnt[top - 1]);   // light tasks: quarter children one size below
  }
  assert(static_cast<int>(s.tasks.size()) == pk_task_entries(bits));
  // per-warp lists of the warp-local stages (bits >= 6: a quarter of the transform holds whole leaves)
  if (bits >= 6) {
    const int nlocal = pk_local_phases(bits), shift = bits - 2;
    s.local.assign(static_cast<size_t>(pk_local_header(bits)), 0);
    auto add_lists = [&](int stage, const std::vector<uint16_t> &mainl, const std::vector<uint16_t> &lightl, bool spread) {
      for (int w = 0; w < kPkWarps; ++w) {
        std::vector<uint16_t> a, b;
        for (uint16_t o : mainl) if ((o >> shift) == w) a.push_back(o);
        for (uint16_t o : lightl) if ((o >> shift) == w) b.push_back(o);
        if (spread) { a = spread_with_holes(a); b = spread_with_holes(b); }
        uint16_t *hd = &s.local[static_cast<size_t>(4 * (stage * kPkWarps + w))];
        hd[0] = static_cast<uint16_t>(s.local.size()); hd[1] = static_cast<uint16_t>(a.size());
        s.local.insert(s.local.end(), a.begin(), a.end());
        hd = &s.local[static_cast<size_t>(4 * (stage * kPkWarps + w))];
        hd[2] = static_cast<uint16_t>(s.local.size()); hd[3] = static_cast<uint16_t>(b.size());
        s.local.insert(s.local.end(), b.begin(), b.end());
      }
    };
    add_lists(0, h.leaf16_off, h.leaf8_off, true);
    for (int ph = 0; ph < nlocal; ++ph) {
      const int lg = pl.lg[ph], depth = pl.depth[ph];
      const int base = pk_phase_base(bits, ph), nm = pk_phase_main(bits, lg, depth), nl = pk_phase_light(bits, lg, depth);
      const std::vector<uint16_t> mainl(s.tasks.begin() + base, s.tasks.begin() + base + nm);
      const std::vector<uint16_t> lightl(s.tasks.begin() + base + nm, s.tasks.begin() + base + nm + nl);
      // a local task must stay inside the quarter its warp owns
      for (uint16_t o : mainl) { assert((o >> shift) == ((o + ((4 << (depth - 1)) - 1) * (1 << (lg - 2))) >> shift)); (void)o; }
      for (uint16_t o : lightl) { assert((o >> shift) == ((o + ((4 << (depth - 2)) - 1) * (1 << (lg - 2))) >> shift)); (void)o; }
      add_lists(1 + ph, mainl, lightl, false);
    }
    assert(static_cast<int>(s.local.size()) >= pk_local_entries(bits));   // + the holes of the leaf lists
  }
  for (int inv = 0; inv < 2; ++inv) {
    s.perm[inv].assign(static_cast<size_t>(m), 0);
    for (int p = 0; p < m; ++p) {
      const int natural = (-split_radix_index(p, m, inv)) & (m - 1);
      s.perm[inv][static_cast<size_t>(natural)] = static_cast<uint16_t>(pk_slot(p));
    }
  }
  return s;
}

template <class T> std::vector<T> cos_quarter_table(int bits)
{
  const int m = 1 << bits;
  const double freq = 2 * kPi / m;
  std::vector<T> t(static_cast<size_t>(m / 4) + 1);
  for (int i = 0; i <= m / 4; ++i) t[i] = static_cast<T>(std::cos(i * freq));
  return t;
}

template <class T> std::vector<T> twiddle_pyramid(const CfftHostSched &s)
{
  std::vector<T> pyr(static_cast<size_t>(s.pyr_len));
  for (int lg = 5; lg <= s.bits; ++lg) {
    const std::vector<T> row = cos_quarter_table<T>(lg);
    for (size_t k = 0; k < row.size(); ++k) pyr[s.pyr_off[lg] + k] = row[k];
  }
  return pyr;
}

template <class T> void leaf_constants(T &sqrthalf, T &c16_1, T &c16_3)
{
  const std::vector<T> c16 = cos_quarter_table<T>(4);
  sqrthalf = static_cast<T>(kSqrtHalf);
  c16_1 = c16[1];
  c16_3 = c16[3];
}

template std::vector<float> cos_quarter_table<float>(int);
template std::vector<double> cos_quarter_table<double>(int);
template std::vector<float> twiddle_pyramid<float>(const CfftHostSched &);
template std::vector<double> twiddle_pyramid<double>(const CfftHostSched &);
template void leaf_constants<float>(float &, float &, float &);
template void leaf_constants<double>(double &, double &, double &);

// ---------------------------------------------------------------------------------------------------
// fp64 DFT-stage kernel tables
// ---------------------------------------------------------------------------------------------------
std::vector<double> unit_circle(int n, int count, int sign)
{
  std::vector<double> t(2 * static_cast<size_t>(count));
  const double step = 2.0 * M_PI / n;
  for (int k = 0; k < count; ++k) {
    // reduce k to the first octant of the n-point circle
    int r = k % n, oct = 0;
    double c, s;
    if (n % 8 == 0) {
      const int e = n / 8;
      oct = r / e;
      const int rr = r - oct * e;
      const int m = (oct & 1) ? e - rr : rr;               // mirrored in odd octants
      const double cc = std::cos(step * m), ss = std::sin(step * m);
      switch (oct) {
        case 0: c = cc; s = ss; break;
        case 1: c = ss; s = cc; break;
        case 2: c = -ss; s = cc; break;
        case 3: c = -cc; s = ss; break;
        case 4: c = -cc; s = -ss; break;
        case 5: c = -ss; s = -cc; break;
        case 6: c = ss; s = -cc; break;
        default: c = cc; s = -ss; break;
      }
    } else { c = std::cos(step * r); s = std::sin(step * r); }
    t[2 * static_cast<size_t>(k)] = c;
    t[2 * static_cast<size_t>(k) + 1] = sign < 0 ? -s : s;
  }
  return t;
}

std::vector<double> real_spectrum(const std::vector<double> &taps, double scale)
{
  const int n = static_cast<int>(taps.size());
  int bits = 0;
  while ((1 << bits) < n) ++bits;
  // plain iterative radix-2 decimation-in-time transform of the real sequence taken as complex
  std::vector<double> re(static_cast<size_t>(n)), im(static_cast<size_t>(n), 0.0);
  for (int i = 0; i < n; ++i) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((i >> b) & 1) << (bits - 1 - b);
    re[static_cast<size_t>(r)] = taps[static_cast<size_t>(i)];
  }
  const std::vector<double> w = unit_circle(n, n / 2 > 0 ? n / 2 : 1, -1);
  for (int len = 2; len <= n; len <<= 1) {
    const int half = len >> 1, stride = n / len;
    for (int base = 0; base < n; base += len)
      for (int j = 0; j < half; ++j) {
        const double wr = w[2 * static_cast<size_t>(j * stride)], wi = w[2 * static_cast<size_t>(j * stride) + 1];
        const size_t a = static_cast<size_t>(base + j), b = a + static_cast<size_t>(half);
        const double tr = re[b] * wr - im[b] * wi, ti = re[b] * wi + im[b] * wr;
        re[b] = re[a] - tr; im[b] = im[a] - ti;
        re[a] += tr; im[a] += ti;
      }
  }
  std::vector<double> out(2 * (static_cast<size_t>(n) / 2 + 1));
  for (int k = 0; k <= n / 2; ++k) { out[2 * static_cast<size_t>(k)] = scale * re[static_cast<size_t>(k)]; out[2 * static_cast<size_t>(k) + 1] = scale * im[static_cast<size_t>(k)]; }
  return out;
}

}  // namespace b200rate
