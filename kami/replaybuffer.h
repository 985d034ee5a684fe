// kami::ReplayBuffer -- host-side sink of finished trajectories, same interface as the reference
// (kami/replaybuffer.h:10-92): a mutex-guarded ring of (obs[obsize], mcts[psize], result) rows,
// uniform-with-replacement batch selection with rand().  Not on the device hot path: the
// kernels keep samples compact on the GPU and kb_pool_drain_samples expands them into `add`.
#pragma once
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace kami {
class ReplayBuffer {
   public:
    ReplayBuffer(int obsize, int psize, int bufsize)
        : obsize(obsize), psize(psize), bufsize(bufsize), inputs((size_t)obsize * bufsize), mcts((size_t)psize * bufsize), results(bufsize) {}

    void clear() {
        total = 0;
        write_index = 0;
    }
    void add(float* input, float* policy, float result) {
        std::lock_guard<std::mutex> g(mut);
        memcpy(&inputs[(size_t)write_index * obsize], input, sizeof(float) * obsize);
        memcpy(&mcts[(size_t)write_index * psize], policy, sizeof(float) * psize);
        results[write_index] = result;
        write_index = (write_index + 1) % bufsize;
        ++total;
    }
    int size() { return bufsize; }
    long count() { return total; }
    void select_batch(float* dst_input, float* dst_mcts, float* dst_result, int n) {
        std::lock_guard<std::mutex> g(mut);
        for (int i = 0; i < n; ++i) {
            int src = rand() % bufsize;
            memcpy(dst_input + (size_t)i * obsize, &inputs[(size_t)src * obsize], sizeof(float) * obsize);
            memcpy(dst_mcts + (size_t)i * psize, &mcts[(size_t)src * psize], sizeof(float) * psize);
            dst_result[i] = results[src];
        }
    }

   private:
    int obsize, psize, bufsize;
    std::mutex mut;
    std::vector<float> inputs, mcts, results;
    int write_index = 0;
    long total = 0;
};
}  // namespace kami
