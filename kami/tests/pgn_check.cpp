// Env::pgn() check driver (tests/test_dropin.py): reads games from stdin, one per line as
// "<n> <action_1> ... <action_n>", replays each on a kami::Env and prints Env::pgn() on one line.
// Not a reference test; built by `make -C kami dropin`.
#include <iostream>
#include <sstream>
#include <string>

#include "../env.h"

int main() {
    using namespace kami;
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream in(line);
        int n = 0;
        if (!(in >> n)) continue;
        Env env;
        for (int i = 0; i < n; ++i) {
            int a;
            in >> a;
            env.push(a);
        }
        std::cout << "PGN " << env.pgn() << std::endl;
    }
    return 0;
}
