// End-to-end smoke of the whole loop on small settings: self-play fills the replay buffer on the device,
// the training thread clones the model, trains it (CUDA training step), runs the arena and accepts or rejects.
// Not a reference test (test/selfplay.cpp just sleeps forever); built by `make -C kami dropin`.
#include <chrono>
#include <iostream>
#include <thread>

#include "../selfplay.h"

int main() {
    using namespace kami;
    options::setInt("filters", 64);
    options::setInt("residuals", 1);
    options::setInt("selfplay_batch", 32);
    options::setInt("selfplay_nodes", 6);
    options::setInt("replaybuffer_size", 96);
    options::setInt("training_sample_pct", 50);
    options::setInt("training_epochs", 1);
    options::setInt("training_batchsize", 8);
    options::setInt("training_mlr", 2);
    options::setInt("evaluate_batch", 2);
    options::setInt("evaluate_games", 2);
    options::setInt("evaluate_nodes", 4);
    options::setInt("evaluate_target_pct", 50);
    options::setInt("inference_threads", 1);
    options::setInt("training_threads", 1);
    options::setStr("model_path", "/tmp/kami_b200_smoke_model.bin");
    NN model(8, 8, NFEATURES, PSIZE);
    Selfplay s(&model);
    s.start();
    const auto t0 = std::chrono::steady_clock::now();
    long seen = -1;
    while (std::chrono::steady_clock::now() - t0 < std::chrono::seconds(150)) {
        std::this_thread::sleep_for(std::chrono::milliseconds(500));
        if (s.get_rbuf().count() != seen) seen = s.get_rbuf().count();
        if (model.get_generation() >= 1) break;
    }
    s.stop();
    std::cout << "replay samples " << s.get_rbuf().count() << " generation " << model.get_generation() << std::endl;
    return 0;
}
