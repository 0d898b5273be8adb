// Host-side check of the persistent GEMM's work-unit order (maskdit_b200/csrc/unit_sched.h, compiled here with g++):
// walks the unit sequence of every CTA group for one problem and prints what tests/test_unit_sched.py asserts on.
//   usage: unit_sched_check num_m_tiles num_n_tiles num_kb splits narrow_last pair_halves grid cg
//   output: "covered <0|1> units <n> makespan_half_tiles <max> min_half_tiles <min> max_gap <waves>"
//     covered            every (slice, m, n) tile was produced exactly once, with a k range tiling [0, num_kb)
//     makespan/min       largest / smallest per-group cost, a full-width tile = 2, a half-width last column tile = 1
//     max_gap            largest distance, in waves (unit index / groups), between the first and the last visit of an m-panel
//                        within a slice - the L2 locality figure (LPT: the whole launch; paired: <= 1)
#include <cstdio>
#include <cstdlib>
#include <map>
#include <tuple>
#include <vector>

#include "unit_sched.h"

int main(int argc, char** argv) {
  if (argc != 9) return 2;
  mdt::GemmParams p = {};
  p.num_m_tiles = atoi(argv[1]), p.num_n_tiles = atoi(argv[2]), p.num_kb = atoi(argv[3]), p.streamk = atoi(argv[4]);
  p.narrow_last = atoi(argv[5]), p.pair_halves = atoi(argv[6]);
  const int grid = atoi(argv[7]), cg = atoi(argv[8]);
  const int groups = grid / cg;
  std::map<std::tuple<int, int, int>, int> seen;  // (slice-start kb0, m, n) -> count
  std::map<std::pair<int, int>, std::pair<long long, long long>> panel_span;  // (kb0, m) -> first/last wave
  long long makespan = 0, mincost = 1LL << 60, total_units = 0;
  bool ok = true;
  for (int g = 0; g < groups; ++g) {
    mdt::UnitSched s, s2;
    s.init(p, cg, grid, g * cg);
    s2.init(p, cg, grid, g * cg + (cg - 1));  // the peer CTA of the pair must walk the identical sequence
    long long cost = 0, step = 0;
    while (true) {
      const bool a = s.next(), b = s2.next();
      if (a != b) ok = false;
      if (!a) break;
      if (s.cur_m != s2.cur_m || s.cur_n != s2.cur_n || s.kb0 != s2.kb0 || s.kb1 != s2.kb1) ok = false;
      if (s.m_tile() < 0 || s.m_tile() >= p.num_m_tiles || s.n_tile() < 0 || s.n_tile() >= p.num_n_tiles) ok = false;
      if (s.kb0 < 0 || s.kb1 > p.num_kb || s.kb0 >= s.kb1) ok = false;
      ++seen[std::make_tuple(s.kb0, s.m_tile(), s.n_tile())];
      const bool half = p.narrow_last && s.n_tile() == p.num_n_tiles - 1;
      cost += (half ? 1 : 2) * static_cast<long long>(s.kb1 - s.kb0);
      const long long wave = (s.unit - groups) / groups;  // next() already advanced `unit` by one round
      auto key = std::make_pair(s.kb0, s.m_tile());
      auto it = panel_span.find(key);
      if (it == panel_span.end()) panel_span[key] = {wave, wave};
      else {
        if (wave < it->second.first) it->second.first = wave;
        if (wave > it->second.second) it->second.second = wave;
      }
      ++step;
    }
    total_units += step;
    if (cost > makespan) makespan = cost;
    if (cost < mincost) mincost = cost;
  }
  // coverage: every tile once per slice, and the slices of a tile tile the k range
  std::map<std::pair<int, int>, long long> ksum;
  for (auto& kv : seen) {
    if (kv.second != 1) ok = false;
    ksum[{std::get<1>(kv.first), std::get<2>(kv.first)}] += 1;
  }
  if (static_cast<long long>(ksum.size()) != static_cast<long long>(p.num_m_tiles) * p.num_n_tiles) ok = false;
  for (auto& kv : ksum)
    if (kv.second != p.streamk) ok = false;
  long long max_gap = 0;
  for (auto& kv : panel_span)
    if (kv.second.second - kv.second.first > max_gap) max_gap = kv.second.second - kv.second.first;
  // costs are in (half tiles x k-blocks); report per k-block of one slice-free tile
  printf("covered %d units %lld makespan_half_tiles %.3f min_half_tiles %.3f max_gap %lld\n", ok ? 1 : 0, total_units,
         static_cast<double>(makespan) / p.num_kb, static_cast<double>(mincost) / p.num_kb, max_gap);
  return ok ? 0 : 1;
}
