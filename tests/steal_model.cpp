// Host model of the ticket-board protocol (defer_b200/csrc/steal_board.h): the SAME claim / complete / arm code the
// CUDA kernel runs, driven by std::threads standing in for CTAs.  Checks, over many lanes, runs and random op lists:
//   * every tile of every op of every run is processed exactly once;
//   * a tile of op o is never claimed before every tile of op o-1 of that lane and run is complete, and the data
//     those tiles wrote is visible to the claimer (plain loads: ThreadSanitizer flags a missing happens-before edge);
//   * re-arming a lane while CTAs of other lanes keep polling it never hands out work of the wrong run;
//   * a lane's workers leave exactly when their own lane's run is complete.
// usage: steal_model <lanes> <workers_per_lane> <runs> <seed>      exit code 0 = all invariants held
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "../defer_b200/csrc/steal_board.h"

struct Op {
  int m_tiles, n_tiles;
};
using Board = defer::LaneBoardT<Op>;

struct Run {
  std::vector<Op> ops;
  std::vector<std::vector<std::atomic<int>>> processed;   // [op][tile] times processed
  std::vector<std::vector<int>> data;                     // [op][tile] payload written by the tile (plain memory)
  std::vector<std::atomic<int>> completed;                // [op] tiles completed
};

static std::atomic<int> g_errors{0};
#define CHECK(cond, ...)                                  \
  do {                                                    \
    if (!(cond)) {                                        \
      if (g_errors.fetch_add(1) < 20) {                   \
        std::fprintf(stderr, "INVARIANT VIOLATED: ");     \
        std::fprintf(stderr, __VA_ARGS__);                \
        std::fprintf(stderr, "\n");                       \
      }                                                   \
    }                                                     \
  } while (0)

int main(int argc, char** argv) {
  const int lanes = argc > 1 ? std::atoi(argv[1]) : 6;
  const int workers = argc > 2 ? std::atoi(argv[2]) : 3;
  const int runs = argc > 3 ? std::atoi(argv[3]) : 200;
  const unsigned seed = argc > 4 ? (unsigned)std::atoi(argv[4]) : 1u;
  std::vector<Board> boards(lanes);
  for (auto& b : boards) {
    b.ticket = 0; b.done = 0; b.target_epoch = 0; b.ops = nullptr; b.n_ops = 0;
  }
  // all runs of all lanes up front (no tracking state is ever reset: a stale claim shows up as a wrong count)
  std::vector<std::vector<Run>> plan(lanes);
  std::mt19937 rng(seed);
  for (int l = 0; l < lanes; ++l) {
    plan[l] = std::vector<Run>(runs);
    for (auto& r : plan[l]) {
      const int n_ops = 1 + (int)(rng() % 7);
      r.ops.resize(n_ops);
      r.processed = std::vector<std::vector<std::atomic<int>>>(n_ops);
      r.data.resize(n_ops);
      r.completed = std::vector<std::atomic<int>>(n_ops);
      for (int o = 0; o < n_ops; ++o) {
        r.ops[o].m_tiles = 1 + (int)(rng() % 4);
        r.ops[o].n_tiles = 1 + (int)(rng() % 3);
        const int nt = r.ops[o].m_tiles * r.ops[o].n_tiles;
        r.processed[o] = std::vector<std::atomic<int>>(nt);
        r.data[o].assign(nt, -1);
        for (auto& a : r.processed[o]) a.store(0);
        r.completed[o].store(0);
      }
    }
  }
  // which run a lane is in (for locating the tracking state of a claim): epoch e <-> run e-1 of that lane
  auto worker = [&](int lane, int run_index, unsigned wseed) {
    std::mt19937 wr(wseed);
    Board* mine = &boards[lane];
    const unsigned my_epoch = defer::sb_ld_relaxed(&mine->target_epoch);
    CHECK((int)my_epoch == run_index + 1, "lane %d worker sees epoch %u in run %d", lane, my_epoch, run_index);
    while (true) {
      if (defer::steal_lane_done(mine, my_epoch)) break;
      defer::StealClaim cl;
      const Op* op = nullptr;
      if (!defer::steal_try_claim(boards.data(), lanes, lane, &cl, &op)) {
        std::this_thread::yield();
        continue;
      }
      // the claimed lane's run = its current epoch - 1 (read after the claim: the lane cannot be re-armed while one of
      // its tiles is outstanding)
      const unsigned e = defer::sb_ld_relaxed(&boards[cl.lane].target_epoch);
      CHECK(e >= 1 && (int)e <= runs, "claim in epoch %u", e);
      Run& r = plan[cl.lane][e - 1];
      CHECK(cl.op < (int)r.ops.size(), "lane %d run %u: op %d out of range", cl.lane, e - 1, cl.op);
      if (cl.op >= (int)r.ops.size()) continue;
      CHECK(op == &r.ops[cl.op], "lane %d run %u: descriptor of another run", cl.lane, e - 1);
      const int nt = r.ops[cl.op].m_tiles * r.ops[cl.op].n_tiles;
      CHECK(cl.tile < nt, "tile %d of %d", cl.tile, nt);
      if (cl.tile >= nt) continue;
      if (cl.op > 0) {
        const int pt = r.ops[cl.op - 1].m_tiles * r.ops[cl.op - 1].n_tiles;
        CHECK(r.completed[cl.op - 1].load() == pt, "lane %d run %u: op %d claimed with op %d at %d/%d tiles", cl.lane, e - 1,
              cl.op, cl.op - 1, r.completed[cl.op - 1].load(), pt);
        for (int t = 0; t < pt; ++t)     // plain loads of what the previous op's tiles wrote
          CHECK(r.data[cl.op - 1][t] == (int)e, "lane %d run %u: stale data of op %d tile %d", cl.lane, e - 1, cl.op - 1, t);
      }
      const int before = r.processed[cl.op][cl.tile].fetch_add(1);
      CHECK(before == 0, "lane %d run %u op %d tile %d processed twice", cl.lane, e - 1, cl.op, cl.tile);
      for (volatile int spin = (int)(wr() % 200); spin > 0; --spin) {}
      r.data[cl.op][cl.tile] = (int)e;   // the tile's "stores"
      r.completed[cl.op].fetch_add(1);
      defer::steal_complete(&boards[cl.lane], cl.op, (unsigned)nt);
    }
  };
  // one orchestrator per lane = the lane's stream: arm -> lane kernel (workers) -> next run
  std::vector<std::thread> streams;
  for (int l = 0; l < lanes; ++l) {
    streams.emplace_back([&, l]() {
      for (int r = 0; r < runs; ++r) {
        defer::steal_arm(&boards[l], plan[l][r].ops.data(), (int)plan[l][r].ops.size());
        std::vector<std::thread> ctas;
        for (int w = 0; w < workers; ++w) ctas.emplace_back(worker, l, r, seed * 7919u + l * 131u + r * 17u + w);
        for (auto& t : ctas) t.join();
        // the lane's kernel has left: its run must be complete
        for (size_t o = 0; o < plan[l][r].ops.size(); ++o) {
          const int nt = plan[l][r].ops[o].m_tiles * plan[l][r].ops[o].n_tiles;
          CHECK(plan[l][r].completed[o].load() == nt, "lane %d run %d left with op %zu at %d/%d", l, r, o,
                plan[l][r].completed[o].load(), nt);
        }
      }
    });
  }
  for (auto& t : streams) t.join();
  long tiles = 0;
  for (int l = 0; l < lanes; ++l)
    for (auto& r : plan[l])
      for (size_t o = 0; o < r.ops.size(); ++o)
        for (auto& a : r.processed[o]) {
          CHECK(a.load() == 1, "lane %d: a tile was processed %d times", l, a.load());
          ++tiles;
        }
  std::printf("%ld tiles, %d lanes x %d workers x %d runs, errors %d\n", tiles, lanes, workers, runs, g_errors.load());
  return g_errors.load() ? 1 : 0;
}
