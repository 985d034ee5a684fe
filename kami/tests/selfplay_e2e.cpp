// End-to-end throughput of the product's own public API: kami::Selfplay (kami/selfplay.h) exactly as kami.cpp drives
// it -- inference threads on device-resident pools (kb_pool_step), finished games drained into the host ReplayBuffer
// (kb_pool_drain_samples -> ReplayBuffer::add), training threads off.  Prints ONE JSON line: NN evals/s, positions/s,
// replay rows/s over a wall-clock window after a warm-up, and the bytes that crossed PCIe per 1024-eval step.
//   selfplay_e2e [seconds=6] [devices=1] [trees_per_thread=1024] [filters=64] [residuals=2] [nodes=1024]
// Not a reference test; built by `make -C kami dropin`.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <thread>

#include "../selfplay.h"

int main(int argc, char** argv) {
    using namespace kami;
    const double seconds = argc > 1 ? atof(argv[1]) : 6.0;
    const int devices = argc > 2 ? atoi(argv[2]) : 1;
    const int trees = argc > 3 ? atoi(argv[3]) : 1024;
    const int filters = argc > 4 ? atoi(argv[4]) : 64;
    const int residuals = argc > 5 ? atoi(argv[5]) : 2;
    const int nodes = argc > 6 ? atoi(argv[6]) : 1024;
    // options.def.yml values of the keys the hot path reads
    options::setInt("filters", filters);
    options::setInt("residuals", residuals);
    options::setInt("selfplay_batch", trees);
    options::setInt("selfplay_nodes", nodes);
    options::setFloat("cpuct", 1.5f);
    options::setInt("unvisited_node_value_pct", 50);
    options::setInt("bootstrap_weight", 20);
    options::setFloat("mcts_noise_weight", 0.05f);
    options::setFloat("selfplay_alpha_initial", 1.0f);
    options::setFloat("selfplay_alpha_decay", 0.95f);
    options::setFloat("selfplay_alpha_final", 0.5f);
    options::setFloat("selfplay_alpha_cutoff", 20.0f);
    options::setInt("replaybuffer_size", 1 << 15);
    options::setInt("inference_threads", devices);
    options::setInt("training_threads", 0);
    options::setInt("b200_devices", devices);
    options::setInt("b200_node_capacity", 1 << 19);
    if (kb_init(0) != KB_OK) {
        fprintf(stderr, "%s\n", kb_last_error());
        return 1;
    }
    NN model(8, 8, NFEATURES, PSIZE);
    Selfplay s(&model);
    s.start();
    // warm-up: let the trees reach steady state (first moves made, games of different lengths in flight)
    std::this_thread::sleep_for(std::chrono::milliseconds(2500));
    const auto t0 = std::chrono::steady_clock::now();
    const unsigned long long e0 = s.evals(), m0 = s.moves();
    const long r0 = s.get_rbuf().count();
    std::this_thread::sleep_for(std::chrono::milliseconds((long)(seconds * 1000)));
    const unsigned long long e1 = s.evals(), m1 = s.moves();
    const long r1 = s.get_rbuf().count();
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    s.stop();
    const double evals = (double)(e1 - e0), rows = (double)(r1 - r0);
    // device -> host per replay row: 80 B position + sparse visit list + z (624 B); the dense rows are expanded on the host.
    // host -> device: nothing (weights are resident; no per-step input comes from the host)
    printf("{\"api\": \"kami::Selfplay (kami/selfplay.h) start/stop, %d inference thread(s) on %d GPU(s)\", \"evals_per_sec\": %.1f, "
           "\"positions_per_sec\": %.1f, \"replay_rows_per_sec\": %.1f, \"seconds\": %.2f, \"trees_per_thread\": %d, \"filters\": %d, "
           "\"residuals\": %d, \"selfplay_nodes\": %d, \"d2h_bytes_per_1024_evals\": %.1f, \"h2d_bytes_per_1024_evals\": 0}\n",
           devices, devices, evals / dt, (double)(m1 - m0) / dt, rows / dt, dt, trees, filters, residuals, nodes,
           evals > 0 ? rows * 624.0 / (evals / 1024.0) : 0.0);
    return 0;
}
