// Smoke driver for kami::eval (not a reference test: the reference ships none for the arena).
// Two random-init networks, tiny search; prints the verdict.  Built by `make -C kami dropin`.
#include <cstdlib>
#include <iostream>

#include "../evaluate.h"

int main() {
    using namespace kami;
    options::setInt("filters", 64);
    options::setInt("residuals", 1);
    options::setInt("evaluate_batch", 2);
    options::setInt("evaluate_games", 3);
    options::setInt("evaluate_nodes", 12);
    options::setInt("evaluate_target_pct", 54);
    srand(7);
    NN current(8, 8, NFEATURES, PSIZE);
    NN candidate(&current);
    // the candidate must be a newer generation or eval() bails out (evaluate.cpp:54-60)
    const bool stale = eval(&current, &candidate, 0);
    std::cout << "stale verdict " << stale << std::endl;
    float in[8 * 8 * NFEATURES] = {0}, pi[PSIZE] = {0}, z[8] = {0};
    options::setInt("training_epochs", 1);
    options::setInt("training_batchsize", 1);
    pi[0] = 1.0f;
    candidate.train(1, in, pi, z);  // generation 1
    const bool ok = eval(&current, &candidate, 0);
    std::cout << "arena verdict " << ok << std::endl;
    return 0;
}
