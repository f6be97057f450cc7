// mppi_resident_host.h — host side of the resident-mode protocol (device side: mppi_resident.cuh).
//
// Pure C++ (no CUDA): everything the protocol needs from the device runtime goes through ResidentBackend, so the
// same code runs in the library (backend = cooperative launch / stream synchronise / stream query) and in
// tests/test_resident_protocol.py, where the "grid" is a host thread that follows the kernel's side of the protocol.
//
// host_box (pinned host memory, 8-byte words):
//   [0, 32)   the command record: word w = payload32 | seq32 << 32;  w0 = flags (bit 0 shift, bit 1 stop),
//             w1 / w2 = Philox counter lo / hi, w3.. = the start state (f32: one word per value, f64: lo, hi),
//             then, for controllers sharded over several GPUs, the exchange epoch lo / hi
//   [32]      sequence number of the last command whose device-side results are complete
//   [33]      exit word: launch generation << 32 | reason (1 idle clock, 2 stop record, 3 lost the finisher)
//   [64, ..)  the action: payload32 | seq32 << 32 per word (f64: lo, hi)
// Every word validates itself, so neither side needs a fence or a particular store order.
#pragma once

#include <stdint.h>
#include <string.h>

#include <chrono>

namespace mppi {

enum { RES_BOX_RECORD = 0, RES_BOX_DONE = 32, RES_BOX_EXIT = 33, RES_BOX_ACTION = 64 };
enum { RES_CMD_SHIFT = 1u, RES_CMD_STOP = 2u };
enum { RES_OK = 0, RES_ERR_BAD_ARG = -1, RES_ERR_TIMEOUT = -6 };      // values of MppiStatus

struct ResidentBackend {
    void* ctx;
    // put a grid on the device that waits for record seq_start + 1; 0 or a (negative) status
    int (*launch)(void* ctx, uint64_t seed, uint64_t offset_pred, int shift_pred, uint64_t seq_start, uint32_t gen);
    // wait until no grid of this controller is on the device any more
    int (*drain)(void* ctx);
    // 0 while the device is healthy (grid running, or gone without an error), else a (negative) status
    int (*health)(void* ctx);
};

struct Resident {
    int armed = 0, launched = 0;
    uint32_t gen = 0;
    uint64_t seq = 0;          // last sequence number handed out (commands and stop records)
    uint64_t cmd_seq = 0;      // sequence number of the last COMMAND (what box[RES_BOX_DONE] converges to)
    uint64_t seed = 0;
    uint64_t launches = 0;     // grids launched so far: the first command and every wake-up after an idle exit
    volatile uint64_t* box = nullptr;
    int nx = 0, n_action = 0, is_double = 0;
    int xchg = 0;              // sharded controller: the record also carries the exchange epoch
    int timeout_s = 10;
    ResidentBackend be{};
};

inline void res_cpu_relax() {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
}

inline int res_record_words(const Resident& r) { return 3 + r.nx * (r.is_double ? 2 : 1) + (r.xchg ? 2 : 0); }
inline int res_box_words(int n_action, int is_double) { return RES_BOX_ACTION + n_action * (is_double ? 2 : 1); }

inline bool res_exited(const Resident& r) { return (uint32_t)(r.box[RES_BOX_EXIT] >> 32) == r.gen; }

inline int res_launch(Resident& r, uint64_t seed, uint64_t offset_pred, int shift_pred, uint64_t seq_start) {
    const int rc = r.be.launch(r.be.ctx, seed, offset_pred, shift_pred, seq_start, ++r.gen);
    if (rc) return rc;
    r.launched = 1;
    r.seed = seed;
    ++r.launches;
    return RES_OK;
}

// A stop record (it consumes a sequence number), then wait for the grid to leave.
inline int res_halt(Resident& r) {
    if (!r.launched) return RES_OK;
    const uint64_t seq = ++r.seq, tag = (seq & 0xffffffffull) << 32;
    for (int w = res_record_words(r) - 1; w >= 1; --w) r.box[w] = tag;
    r.box[0] = tag | RES_CMD_STOP;
    r.launched = 0;
    return r.be.drain(r.be.ctx);
}

inline int res_arm(Resident& r, void* host_box, int nx, int n_action, int is_double, const ResidentBackend& be, int xchg = 0) {
    if (host_box == nullptr || nx < 1 || n_action < 1 || 3 + nx * (is_double ? 2 : 1) + (xchg ? 2 : 0) > 32) return RES_ERR_BAD_ARG;
    if (r.launched) {
        const int rc = res_halt(r);
        if (rc) return rc;
    }
    r.box = reinterpret_cast<volatile uint64_t*>(host_box);
    r.nx = nx;
    r.n_action = n_action;
    r.is_double = is_double;
    r.xchg = xchg ? 1 : 0;
    r.be = be;
    // no grid is polling now: clear the box, so that no word left by an earlier controller (whose sequence numbers and
    // launch generations also started at 1) can pass for one of this controller's
    for (int w = 0; w < res_box_words(n_action, is_double); ++w) r.box[w] = 0ull;
    r.armed = 1;
    return RES_OK;
}

// One command: post the record, wait for the action words, copy the action out (controller dtype).
inline int res_command(Resident& r, const double* state, int shift, uint64_t seed, uint64_t offset, void* action_out,
                       uint64_t epoch = 0) {
    if (!r.armed || state == nullptr || action_out == nullptr) return RES_ERR_BAD_ARG;
    int rc;
    if (r.launched && res_exited(r)) r.launched = 0;                             // it left on its idle clock
    if (r.launched && seed != r.seed && (rc = res_halt(r)) != RES_OK) return rc;   // reseeded generator
    if (!r.launched && (rc = res_launch(r, seed, offset, shift, r.seq)) != RES_OK) return rc;
    const uint64_t seq = r.seq + 1, tag = (seq & 0xffffffffull) << 32;
    volatile uint64_t* box = r.box;
    if (r.is_double) {
        for (int i = 0; i < r.nx; ++i) {
            uint64_t bits;
            memcpy(&bits, &state[i], 8);
            box[3 + 2 * i] = tag | (bits & 0xffffffffull);
            box[3 + 2 * i + 1] = tag | (bits >> 32);
        }
    } else {
        for (int i = 0; i < r.nx; ++i) {
            const float f = (float)state[i];
            uint32_t bits;
            memcpy(&bits, &f, 4);
            box[3 + i] = tag | bits;
        }
    }
    if (r.xchg) {
        const int e0 = 3 + r.nx * (r.is_double ? 2 : 1);
        box[e0] = tag | (epoch & 0xffffffffull);
        box[e0 + 1] = tag | (epoch >> 32);
    }
    box[1] = tag | (offset & 0xffffffffull);
    box[2] = tag | (offset >> 32);
    box[0] = tag | (shift ? RES_CMD_SHIFT : 0u);
    r.seq = seq;
    r.cmd_seq = seq;

    const uint64_t want = seq & 0xffffffffull;
    volatile uint64_t* act = box + RES_BOX_ACTION;
    const int nwords = r.n_action * (r.is_double ? 2 : 1);
    uint64_t spins = 0;
    int relaunches = 0;
    const auto t0 = std::chrono::steady_clock::now();
    for (int w = 0; w < nwords; ++w) {
        while ((act[w] >> 32) != want) {
            res_cpu_relax();
            if ((++spins & 0x3FFF) != 0) continue;
            if (res_exited(r) && (act[w] >> 32) != want) {
                // the grid left (idle clock) before it saw this record: wake it up; the record is still in the box.
                // A grid never takes a record after deciding to leave, so the command cannot run twice.
                if (++relaunches > 3) return RES_ERR_TIMEOUT;
                if ((rc = res_launch(r, seed, offset, shift, seq - 1)) != RES_OK) return rc;
            }
            if ((spins & 0xFFFFF) == 0) {
                if ((rc = r.be.health(r.be.ctx)) != RES_OK) return rc;
                if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(r.timeout_s)) return RES_ERR_TIMEOUT;
            }
        }
    }
    if (r.is_double) {
        uint64_t* out = reinterpret_cast<uint64_t*>(action_out);
        for (int i = 0; i < r.n_action; ++i) out[i] = (act[2 * i] & 0xffffffffull) | (act[2 * i + 1] << 32);
    } else {
        uint32_t* out = reinterpret_cast<uint32_t*>(action_out);
        for (int i = 0; i < r.n_action; ++i) out[i] = (uint32_t)act[i];
    }
    return RES_OK;
}

// Wait until everything the last command wrote on the device is complete (the finisher's done word).
inline int res_sync(Resident& r) {
    if (!r.armed || r.cmd_seq == 0) return RES_OK;
    uint64_t spins = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (r.box[RES_BOX_DONE] != r.cmd_seq) {
        res_cpu_relax();
        if ((++spins & 0xFFFFF) == 0) {
            const int rc = r.be.health(r.be.ctx);
            if (rc) return rc;
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(r.timeout_s)) return RES_ERR_TIMEOUT;
        }
    }
    return RES_OK;
}

inline int res_stop(Resident& r) {
    int rc = res_halt(r);
    if (r.armed && rc == RES_OK) rc = r.be.drain(r.be.ctx);     // also covers a grid that is leaving on its idle clock
    r.armed = 0;
    return rc;
}

}  // namespace mppi
