// CPU test harness for the resident-mode protocol (tests/test_resident_protocol.py builds and runs it).
//
// The host side is the library's own code (csrc/mppi_resident_host.h).  The "grid" is a host thread that follows the
// device side of the protocol as csrc/mppi_resident.cuh implements it: poll the record (all words must carry the
// wanted sequence number), check the idle clock only after a failed poll, never take a record after deciding to
// leave, write the action words, then the done word, and an exit word when leaving.  "Stream order" (a relaunched
// grid starts after the previous one is gone) is the backend joining the old thread before it starts a new one.
//
// Test infrastructure only.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "../pytorch_mppi_b200/csrc/mppi_resident_host.h"

using namespace mppi;
using clk = std::chrono::steady_clock;

struct Taken {
    uint64_t seq, offset;
    int shift;
    double state[8];
    uint64_t seed, epoch;
    uint32_t gen;
};

struct Sim {
    volatile uint64_t* box = nullptr;
    int nx = 2, n_action = 1, is_double = 0, xchg = 0;
    uint64_t idle_ns = 50000000ull;
    int compute_us_max = 0;          // emulated command duration: uniform in [0, compute_us_max]
    int done_lag_us = 0;             // extra time between the action words and the done word
    std::thread th;
    std::vector<Taken> log;          // every record a grid took, in order
    int redo = 0;                    // records that contradicted the prediction
    std::vector<int> exit_reasons;
};

static uint64_t now_ns() { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(clk::now().time_since_epoch()).count(); }
static void spin_us(int us) {
    const uint64_t t0 = now_ns();
    while (now_ns() - t0 < (uint64_t)us * 1000ull) res_cpu_relax();
}

// the value an emulated command returns for action element i: any function of everything the record carries
static double expected_action(const double* st, int nx, uint64_t seed, uint64_t offset, int shift, int i, int is_double) {
    double s = 0.0;
    for (int k = 0; k < nx; ++k) s += (k + 1) * (is_double ? st[k] : (double)(float)st[k]);
    return s + 1e-3 * (double)(offset % 1000003ull) + 0.5 * shift + 0.25 * i + 1e-6 * (double)(seed % 997ull);
}

static void grid_main(Sim* sim, uint64_t seed, uint64_t offset_pred, int shift_pred, uint64_t seq, uint32_t gen) {
    volatile uint64_t* box = sim->box;
    const int nxw = sim->nx * (sim->is_double ? 2 : 1);
    const int nw = 3 + nxw + (sim->xchg ? 2 : 0);
    std::mt19937 rng(gen * 7919u + 13u);
    uint64_t t_idle0 = now_ns();
    int reason = 0;
    for (;;) {
        const uint32_t want = (uint32_t)((seq + 1) & 0xffffffffull);
        uint32_t payload[32];
        for (;;) {
            bool ok = true;
            for (int w = 0; w < nw; ++w) {
                const uint64_t v = box[w];
                if ((uint32_t)(v >> 32) != want) ok = false;
                payload[w] = (uint32_t)v;
            }
            if (ok) break;
            if (now_ns() - t_idle0 > sim->idle_ns) { reason = 1; break; }
        }
        if (reason) break;
        if (payload[0] & RES_CMD_STOP) { reason = 2; break; }
        t_idle0 = now_ns();
        Taken t{};
        t.seq = seq + 1;
        t.shift = (payload[0] & RES_CMD_SHIFT) ? 1 : 0;
        t.offset = (uint64_t)payload[1] | ((uint64_t)payload[2] << 32);
        t.seed = seed;
        t.gen = gen;
        for (int i = 0; i < sim->nx; ++i) {
            if (sim->is_double) {
                const uint64_t bits = (uint64_t)payload[3 + 2 * i] | ((uint64_t)payload[3 + 2 * i + 1] << 32);
                memcpy(&t.state[i], &bits, 8);
            } else {
                float f;
                memcpy(&f, &payload[3 + i], 4);
                t.state[i] = f;
            }
        }
        if (sim->xchg) t.epoch = (uint64_t)payload[3 + nxw] | ((uint64_t)payload[4 + nxw] << 32);
        if (t.offset != offset_pred || t.shift != shift_pred) ++sim->redo;
        sim->log.push_back(t);
        if (sim->compute_us_max > 0) spin_us((int)(rng() % (unsigned)(sim->compute_us_max + 1)));
        const uint64_t tag = (uint64_t)want << 32;
        for (int i = 0; i < sim->n_action; ++i) {
            const double a = expected_action(t.state, sim->nx, seed, t.offset, t.shift, i, sim->is_double);
            if (sim->is_double) {
                uint64_t bits;
                memcpy(&bits, &a, 8);
                box[RES_BOX_ACTION + 2 * i] = tag | (bits & 0xffffffffull);
                box[RES_BOX_ACTION + 2 * i + 1] = tag | (bits >> 32);
            } else {
                const float f = (float)a;
                uint32_t bits;
                memcpy(&bits, &f, 4);
                box[RES_BOX_ACTION + i] = tag | bits;
            }
        }
        if (sim->done_lag_us > 0) spin_us(sim->done_lag_us);
        box[RES_BOX_DONE] = seq + 1;
        ++seq;
        offset_pred = t.offset + 8;          // the tests advance the counter by 8 per command
        shift_pred = t.shift;
    }
    sim->exit_reasons.push_back(reason);
    box[RES_BOX_EXIT] = ((uint64_t)gen << 32) | (uint64_t)reason;
}

static int be_launch(void* ctx, uint64_t seed, uint64_t offset_pred, int shift_pred, uint64_t seq_start, uint32_t gen) {
    Sim* sim = reinterpret_cast<Sim*>(ctx);
    if (sim->th.joinable()) sim->th.join();          // stream order
    sim->th = std::thread(grid_main, sim, seed, offset_pred, shift_pred, seq_start, gen);
    return 0;
}
static int be_drain(void* ctx) {
    Sim* sim = reinterpret_cast<Sim*>(ctx);
    if (sim->th.joinable()) sim->th.join();
    return 0;
}
static int be_health(void*) { return 0; }

#define CHECK(cond, ...)                                                    \
    do {                                                                    \
        if (!(cond)) {                                                      \
            fprintf(stderr, "FAIL %s:%d: %s — ", __FILE__, __LINE__, #cond); \
            fprintf(stderr, __VA_ARGS__);                                   \
            fprintf(stderr, "\n");                                          \
            return 1;                                                       \
        }                                                                   \
    } while (0)

static double read_action(const void* buf, int i, int is_double) {
    return is_double ? reinterpret_cast<const double*>(buf)[i] : (double)reinterpret_cast<const float*>(buf)[i];
}

// n commands through a Resident; checks every returned action and the grid's log.  pause_us_max > 0: the host idles a
// random time between commands (to race the grid's idle clock).
static int run_commands(Resident& r, Sim& sim, int n, uint64_t seed, uint64_t& offset, int pause_us_max, std::mt19937& rng,
                        bool flip_shift, uint64_t* first_seq) {
    double st[8];
    unsigned char out[8 * 16];
    const size_t log0 = sim.log.size();
    for (int c = 0; c < n; ++c) {
        for (int k = 0; k < sim.nx; ++k) st[k] = std::sin(0.37 * c + k) * 3.0 + 1e-9 * c;
        const int shift = flip_shift ? ((c % 3) != 2) : 1;
        const int rc = res_command(r, st, shift, seed, offset, out);
        CHECK(rc == 0, "res_command returned %d at command %d", rc, c);
        for (int i = 0; i < sim.n_action; ++i) {
            const double want = expected_action(st, sim.nx, seed, offset, shift, i, sim.is_double);
            const double got = read_action(out, i, sim.is_double);
            const double ref = sim.is_double ? want : (double)(float)want;
            CHECK(got == ref, "command %d action[%d]: got %.17g expected %.17g", c, i, got, ref);
        }
        offset += 8;
        if (pause_us_max > 0) spin_us((int)(rng() % (unsigned)(pause_us_max + 1)));
    }
    CHECK(res_sync(r) == 0, "res_sync");
    CHECK(sim.box[RES_BOX_DONE] == r.cmd_seq, "done word %llu != last command %llu", (unsigned long long)sim.box[RES_BOX_DONE],
          (unsigned long long)r.cmd_seq);
    // exactly once, in order
    CHECK(sim.log.size() - log0 == (size_t)n, "grid took %zu records for %d commands", sim.log.size() - log0, n);
    for (size_t i = log0 + 1; i < sim.log.size(); ++i)
        CHECK(sim.log[i].seq > sim.log[i - 1].seq, "records out of order at %zu", i);
    if (first_seq != nullptr) *first_seq = sim.log[log0].seq;
    return 0;
}

int main() {
    std::mt19937 rng(12345);
    alignas(64) static uint64_t boxmem[RES_BOX_ACTION + 64];
    ResidentBackend be{nullptr, be_launch, be_drain, be_health};

    // 1. f32, back to back: one launch serves everything; predictions hold (no redo)
    {
        Sim sim;
        sim.box = boxmem;
        be.ctx = &sim;
        Resident r;
        CHECK(res_arm(r, boxmem, 2, 1, 0, be) == 0, "arm");
        uint64_t off = 100;
        if (run_commands(r, sim, 2000, 7, off, 0, rng, false, nullptr)) return 1;
        CHECK(r.launches == 1, "launches = %llu", (unsigned long long)r.launches);
        CHECK(sim.redo == 0, "redo = %d", sim.redo);
        CHECK(res_stop(r) == 0, "stop");
        CHECK(sim.exit_reasons.size() == 1 && sim.exit_reasons[0] == 2, "exit reason");
        printf("1 ok: 2000 commands, 1 launch\n");
    }
    // 2. f64, nx = 3, four action values, alternating shift flag: every third command contradicts the prediction
    {
        Sim sim;
        sim.box = boxmem;
        sim.nx = 3;
        sim.n_action = 4;
        sim.is_double = 1;
        sim.compute_us_max = 5;
        sim.done_lag_us = 3;
        be.ctx = &sim;
        Resident r;
        CHECK(res_arm(r, boxmem, 3, 4, 1, be) == 0, "arm");
        uint64_t off = (1ull << 40) + 5;          // exercises the high counter word
        if (run_commands(r, sim, 600, 99, off, 0, rng, true, nullptr)) return 1;
        CHECK(sim.redo == 400 - 1 || sim.redo == 400, "redo = %d", sim.redo);     // two flips per period of three
        CHECK(res_stop(r) == 0, "stop");
        printf("2 ok: f64 records, %d mispredictions handled\n", sim.redo);
    }
    // 3. the idle clock races the host: pauses around idle_ns; nothing lost, nothing run twice, grids relaunched
    {
        Sim sim;
        sim.box = boxmem;
        sim.idle_ns = 30000;
        sim.compute_us_max = 8;
        be.ctx = &sim;
        Resident r;
        CHECK(res_arm(r, boxmem, 2, 1, 0, be) == 0, "arm");
        uint64_t off = 0;
        if (run_commands(r, sim, 4000, 3, off, 60, rng, false, nullptr)) return 1;
        CHECK(r.launches > 10, "launches = %llu (the idle clock never fired?)", (unsigned long long)r.launches);
        CHECK(res_stop(r) == 0, "stop");
        int idle_exits = 0;
        for (int e : sim.exit_reasons) idle_exits += e == 1;
        CHECK(idle_exits + 1 >= (int)r.launches, "exits");
        printf("3 ok: 4000 commands against a 30 us idle clock, %llu launches, %d idle exits\n", (unsigned long long)r.launches,
               idle_exits);
    }
    // 4. stop / re-arm on the same controller, then a NEW controller on the same (dirty) box, then a reseed
    {
        Sim sim;
        sim.box = boxmem;
        be.ctx = &sim;
        Resident r;
        CHECK(res_arm(r, boxmem, 2, 1, 0, be) == 0, "arm");
        uint64_t off = 0, s0 = 0, s1 = 0;
        if (run_commands(r, sim, 5, 1, off, 0, rng, false, &s0)) return 1;
        CHECK(res_stop(r) == 0, "stop");
        CHECK(res_arm(r, boxmem, 2, 1, 0, be) == 0, "re-arm");
        if (run_commands(r, sim, 5, 1, off, 0, rng, false, &s1)) return 1;
        CHECK(s0 == 1 && s1 == 7, "sequence numbers %llu %llu (a stop record consumes one)", (unsigned long long)s0, (unsigned long long)s1);
        CHECK(res_stop(r) == 0, "stop");
        // new controller: sequence numbers and generations start again at 1 on a box full of old words
        Sim sim2;
        sim2.box = boxmem;
        be.ctx = &sim2;
        Resident r2;
        CHECK(res_arm(r2, boxmem, 2, 1, 0, be) == 0, "arm 2");
        uint64_t off2 = 0;
        if (run_commands(r2, sim2, 20, 1, off2, 0, rng, false, nullptr)) return 1;
        CHECK(r2.launches == 1, "stale exit word mistaken for this generation's: %llu launches", (unsigned long long)r2.launches);
        // reseed: the grid's seed is a launch argument, so a new seed means a new grid
        if (run_commands(r2, sim2, 5, 2, off2, 0, rng, false, nullptr)) return 1;
        CHECK(r2.launches == 2 && sim2.log.back().seed == 2, "reseed");
        CHECK(res_stop(r2) == 0, "stop 2");
        printf("4 ok: stop/re-arm, fresh controller on a used box, reseed\n");
    }
    // 5. sharded controller: the record carries the exchange epoch (f64, nx = 8: the longest record, 21 words)
    {
        Sim sim;
        sim.box = boxmem;
        sim.nx = 8;
        sim.is_double = 1;
        sim.xchg = 1;
        be.ctx = &sim;
        Resident r;
        CHECK(res_arm(r, boxmem, 8, 1, 1, be, 1) == 0, "arm");
        double st[8] = {1, 2, 3, 4, 5, 6, 7, 8};
        double out[1];
        for (int c = 0; c < 50; ++c) {
            const uint64_t epoch = (1ull << 33) + 17 + c;
            CHECK(res_command(r, st, 1, 5, 8ull * c, out, epoch) == 0, "command");
            CHECK(sim.log.back().epoch == epoch, "epoch %llu != %llu", (unsigned long long)sim.log.back().epoch, (unsigned long long)epoch);
            CHECK(out[0] == expected_action(st, 8, 5, 8ull * c, 1, 0, 1), "action");
        }
        CHECK(res_stop(r) == 0, "stop");
        printf("5 ok: exchange epoch in the record\n");
    }
    // 6. argument errors
    {
        Resident r;
        unsigned char out[16];
        double st[2] = {0, 0};
        CHECK(res_command(r, st, 1, 0, 0, out) == RES_ERR_BAD_ARG, "unarmed command");
        CHECK(res_arm(r, nullptr, 2, 1, 0, be) == RES_ERR_BAD_ARG, "null box");
        CHECK(res_arm(r, boxmem, 20, 1, 1, be) == RES_ERR_BAD_ARG, "record too long");
        CHECK(res_sync(r) == 0 && res_stop(r) == 0, "sync/stop on an unarmed controller are no-ops");
        printf("6 ok: argument errors\n");
    }
    printf("ALL OK\n");
    return 0;
}
