// kami::MCTS over the B200 C ABI -- same public surface as the reference's kami/mcts.h:15-349.
// The tree lives in a device-resident node pool (kb_pool of one tree); select / expand / pick /
// push / snapshot are kernel calls.  `root` stays a public Node* whose `children` callers may
// read and sort (test/mcts.cpp:67-71), so a host mirror of the root and its children is
// refreshed after every call that changes them.
#pragma once
#include <cmath>
#include <ctime>
#include <iomanip>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "env.h"
#include "options.h"

namespace kami {
struct Node {
    int n = 0;
    float w = 0.0f;
    float p = 0.0f;
    int action = -1;
    std::vector<Node*> children;
    Node* parent = nullptr;
    float turn = 0.0f;
    float q(float def = 1.0f) { return n > 0 ? w / n : def; }

    void clean() {
        for (auto& c : children) {
            c->clean();
            delete c;
        }
        children.clear();
    }
    std::string debug(Env* e) {
        std::stringstream out;
        float value;
        out << std::setw(6) << e->debug_action(action);
        out << " Visits: " << std::setw(4) << std::to_string(n);
        out << " Average: " << std::to_string(q());
        out << " Policy: " << std::to_string(p);
        out << " Turn: " << std::to_string(turn);
        e->push(action);
        if (e->terminal(&value)) out << " Terminal: " << std::to_string(value);
        e->pop();
        return out.str();
    }
};

class MCTS {
   private:
    Env env;  // host-side twin of the tree's root game (replays the same pushes)
    kb_pool* pool = nullptr;
    int slot = 0;
    bool owns_pool = true;

    void refresh_root() {
        int32_t action[256], n[256];
        float w[256], p[256];
        int k = 0, rn = 0;
        float rw = 0.0f;
        kb_check(kb_tree_root_children(pool, slot, action, n, w, p, 256, &k));
        kb_check(kb_tree_n(pool, slot, &rn));
        kb_check(kb_tree_root_w(pool, slot, &rw));
        root->clean();
        root->n = rn;
        root->w = rw;
        root->turn = -env.turn();
        for (int i = 0; i < k; ++i) {
            Node* c = new Node();
            c->n = n[i];
            c->w = w[i];
            c->p = p[i];
            c->action = action[i];
            c->parent = root;
            c->turn = -root->turn;
            root->children.push_back(c);
        }
    }

   public:
    Node* root = nullptr;

    MCTS() {
        kb_tree_cfg cfg;
        kb_check(kb_tree_default_cfg(&cfg));
        cfg.cpuct = options::getFloat("cpuct", 1.0f);
        cfg.force_expand_unvisited = options::getInt("force_expand_unvisited", 0);
        cfg.unvisited_node_value_pct = options::getInt("unvisited_node_value_pct", 100);
        cfg.bootstrap_weight = options::getInt("bootstrap_weight", 0);
        cfg.bootstrap_window = options::getInt("bootstrap_window", 1600);
        cfg.bootstrap_amp_pct = options::getInt("bootstrap_amp_pct", 75);
        cfg.scale_cpuct_by_actions = options::getInt("scale_cpuct_by_actions", 0);
        cfg.noise_weight = options::getFloat("mcts_noise_weight", 0.05f);
        cfg.seed = (uint64_t)time(NULL);  // mcts.h:99
        cfg.selfplay_nodes = 0;           // moves are made by the caller (pick/push), as in the reference
        kb_check(kb_pool_create(&pool, 1, options::getInt("b200_node_capacity", 1 << 18), &cfg));
        root = new Node();
        root->turn = -env.turn();
    }
    MCTS(const MCTS&) = delete;
    MCTS& operator=(const MCTS&) = delete;
    ~MCTS() {
        if (root) {
            root->clean();
            delete root;
        }
        if (owns_pool) kb_pool_destroy(pool);
    }

    int n() { return root->n; }

   private:
    // The reference's select() leaves its Env AT the leaf until expand() unwinds it (mcts.h:252-254, 320-326) and
    // callers look at it there (evaluate.cpp:80-90 reads get_env().turn()).  The device tree keeps its own
    // positions; this host mirror follows only when somebody asks (get_env() while a leaf is pending).
    bool pending = false;
    int mirrored_plies = 0;
    void mirror_to_leaf() {
        if (!pending || mirrored_plies) return;
        int32_t path[255];
        int depth = 0;
        kb_check(kb_tree_leaf_path(pool, slot, path, 255, &depth));
        for (int i = 0; i < depth; ++i) env.push(path[i]);
        mirrored_plies = depth;
    }
    void mirror_to_root() {
        for (; mirrored_plies > 0; --mirrored_plies) env.pop();
        pending = false;
    }

   public:
    void push(int action) {
        mirror_to_root();
        int rc = kb_tree_push(pool, slot, action);
        if (rc == KB_ERR_NO_CHILD) throw std::runtime_error("no child for action");
        kb_check(rc);
        env.push(action);
        refresh_root();
    }
    int pick(float alpha = 0.0f) {
        if (!root->children.size()) throw std::runtime_error("no children to pick from");
        int action = -1;
        // mcts.h:141-173: rand() is drawn on the temperature path only -- the greedy path must not advance the
        // process-wide stream (evaluate.cpp:22 colours and ReplayBuffer::select_batch read it too)
        const double u = alpha >= 0.1f ? (double)rand() / (double)RAND_MAX : 0.0;
        kb_check(kb_tree_pick(pool, slot, alpha, u, &action));
        return action;
    }
    bool select(float* obs) {
        int need = 0;
        kb_check(kb_tree_select(pool, slot, obs, &need));
        if (!need) refresh_root();  // a terminal leaf was backed up
        pending = need != 0;
        return need != 0;
    }
    void expand(float* policy, float value, bool disable_bootstrap = false) {
        mirror_to_root();
        kb_check(kb_tree_expand(pool, slot, policy, value, disable_bootstrap ? 1 : 0));
        refresh_root();
    }
    Env& get_env() {
        mirror_to_leaf();
        return env;
    }
    void reset() {
        mirrored_plies = 0;
        pending = false;
        kb_check(kb_tree_reset(pool, slot));
        env = Env();
        refresh_root();
    }
    void snapshot(float* pspace) { kb_check(kb_tree_snapshot(pool, slot, pspace)); }
};
}  // namespace kami
