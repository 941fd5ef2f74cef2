// trees.hpp -- host-side validation + re-encoding of the reference's tree arrays.
//
// The reference passes random trees around as pointer-prefix int32 arrays
// (create_grow_tree_arrays main_cpu.go:609-639, tree_to_array :675-693):
//   function node = [op, idx_child1, idx_child2, <child1...>, <child2...>] (absolute indices),
//   terminal      = [id].
// eval_arrays (:755-775) walks that form recursively per fitness case.  The device interpreter
// instead runs a postfix program (left subtree, right subtree, op) with an explicit operand
// stack, so the shim converts once per tree, checking every index on the way: a malformed array
// is reported as an error instead of faulting on the device.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "program.hpp"

namespace gsgp {

constexpr int kNumFunc = 4;   // + - * /  (main_cpu.go:433-440)

// Appends the postfix program of `tree[0..len)` to `out`; returns false and sets err on a
// malformed tree.  max_sp = operand-stack depth the program needs.
struct PrefixFrame { int32_t pos; int32_t next; };

// `stack` is caller-provided scratch (reused across trees: no allocation per tree on the staging path)
inline bool prefix_to_postfix(const int32_t *tree, int32_t len, int32_t n_symbols,
                              std::vector<uint16_t> &out, int32_t &max_sp, const char *&err,
                              std::vector<PrefixFrame> &stack)
{
    stack.clear();
    max_sp = 0;
    int32_t sp = 0, visited = 0;
    if (len <= 0) { err = "empty tree"; return false; }
    auto visit = [&](int32_t pos) -> bool {
        if (pos < 0 || pos >= len) { err = "tree index out of range"; return false; }
        if (++visited > len) { err = "tree has a cycle"; return false; }
        const int32_t id = tree[pos];
        if (id < 0 || id >= n_symbols || id > 0xFFFF) { err = "unknown symbol id in tree"; return false; }
        if (id >= kNumFunc) {
            out.push_back((uint16_t)id);
            if (++sp > max_sp) max_sp = sp;
            return true;
        }
        if (pos + 2 >= len) { err = "truncated function node"; return false; }
        stack.push_back({pos, 1});
        return true;
    };
    if (!visit(0)) return false;
    while (!stack.empty()) {
        PrefixFrame &f = stack.back();
        if (f.next > 2) {
            out.push_back((uint16_t)tree[f.pos]);
            --sp;
            stack.pop_back();
            continue;
        }
        const int32_t child = tree[f.pos + f.next];
        if (child <= f.pos) { err = "child index does not point forward"; return false; }
        ++f.next;
        if (!visit(child)) return false;
    }
    if (sp != 1) { err = "malformed tree"; return false; }
    return true;
}

// Direct compilation of a reference tree array into the interpreter's accumulator program: the same
// program compile_postfix (program.hpp) builds from the postfix form, without the intermediate -- this is
// the per-generation staging path of gsgp_generation (host replay), so it avoids allocations and walks the
// array twice: (1) validate + spill need per function node (post-order), (2) emit in Sethi-Ullman order.
//   need : scratch, len bytes.   frames : scratch stack (grown as needed).
// `out` must have room for one instruction per function node (at least 1).
struct TreeFrame { int32_t pos; uint8_t state, live, first_is_left, pad; };

inline bool compile_prefix_tree(const int32_t *tree, int32_t len, int32_t n_symbols, Ins *out, int &n_out,
                                int &stack_need, std::vector<uint8_t> &need, std::vector<TreeFrame> &frames,
                                const char *&err)
{
    n_out = 0; stack_need = 0;
    if (len <= 0) { err = "empty tree"; return false; }
    auto bad_id = [&](int32_t id) { return id < 0 || id >= n_symbols || id >= kMaxSymbols; };
    if (bad_id(tree[0])) { err = "unknown symbol id in tree"; return false; }
    if (tree[0] >= kNumFunc) {                            // lone terminal (zero_depth trees)
        if (len != 1) { err = "malformed tree"; return false; }
        out[0].code = (uint32_t)A_MOV;
        out[0].ab = (uint32_t)(tree[0] - kNumFunc) << 16;
        n_out = 1;
        return true;
    }
    if (need.size() < (size_t)len) need.resize((size_t)len);
    // pass 1: post-order over the function nodes.  A tree of f function nodes fills 4 f + 1 array words (3 per function
    // node, f + 1 terminals), and the callers reserve len / 4 + 1 instructions per tree: an array whose child indices
    // share nodes (a DAG: forward pointers, no cycle) would expand into more function-node visits than that, so the
    // walk stops at (len - 1) / 4 of them -- no array can make the emission below write past what was reserved.
    int32_t visited = 0, n_func = 0;
    const int32_t max_func = (len - 1) / 4;
    frames.clear();
    frames.push_back(TreeFrame{0, 0, 0, 0, 0});
    while (!frames.empty()) {
        TreeFrame &f = frames.back();
        const int32_t p = f.pos;
        if (f.state == 0) {
            if (p + 2 >= len) { err = "truncated function node"; return false; }
            if (++visited > len) { err = "tree has a cycle"; return false; }
            if (++n_func > max_func) { err = "tree nodes are shared or overlap (more function nodes than the array holds)"; return false; }
        }
        if (f.state < 2) {                                // visit child f.state (0 = left, 1 = right)
            const int32_t c = tree[p + 1 + f.state];
            ++f.state;
            if (c <= p || c >= len) { err = c <= p ? "child index does not point forward" : "tree index out of range"; return false; }
            const int32_t id = tree[c];
            if (bad_id(id)) { err = "unknown symbol id in tree"; return false; }
            if (id < kNumFunc) frames.push_back(TreeFrame{c, 0, 0, 0, 0});
            else if (++visited > len) { err = "tree has a cycle"; return false; }
            continue;
        }
        const int32_t l = tree[p + 1], r = tree[p + 2];
        const bool lt = tree[l] >= kNumFunc, rt = tree[r] >= kNumFunc;
        uint32_t nd;
        if (!lt && !rt) { const uint32_t nl = need[l], nr = need[r]; nd = (nl == nr) ? nl + 1u : (nl > nr ? nl : nr); }
        else nd = lt ? (rt ? 0u : need[r]) : need[l];
        if (nd > 250u) { err = "tree too deep"; return false; }
        need[p] = (uint8_t)nd;
        frames.pop_back();
    }
    // pass 2: emission (mirrors compile_postfix's second pass)
    int spills = 0;
    frames.push_back(TreeFrame{0, 0, 0, 0, 0});
    while (!frames.empty()) {
        TreeFrame &f = frames.back();
        const int32_t p = f.pos, l = tree[p + 1], r = tree[p + 2];
        const uint32_t tok = (uint32_t)tree[p];
        const uint32_t rev = (tok == 1) ? (uint32_t)A_RSUB : (tok == 3) ? (uint32_t)A_RDIV : tok;
        const bool lt = tree[l] >= kNumFunc, rt = tree[r] >= kNumFunc;
        uint32_t code, ab = 0;
        if (f.state == 0) {
            if (lt && rt) {
                code = K_LOADA | tok;
                if (f.live) {
                    code |= K_PUSH | ((uint32_t)spills << kLevelShift);
                    if (++spills > stack_need) stack_need = spills;
                }
                ab = (uint32_t)(tree[l] - kNumFunc) | ((uint32_t)(tree[r] - kNumFunc) << 16);
            } else {
                int32_t child;
                if (!lt && rt) { f.state = 1; child = l; }
                else if (lt && !rt) { f.state = 2; child = r; }
                else { f.first_is_left = need[r] > need[l] ? 0 : 1; f.state = 3; child = f.first_is_left ? l : r; }
                const uint8_t live = f.live;
                frames.push_back(TreeFrame{child, 0, live, 0, 0});
                continue;
            }
        } else if (f.state == 1) { code = tok; ab = (uint32_t)(tree[r] - kNumFunc) << 16; }
        else if (f.state == 2) { code = rev; ab = (uint32_t)(tree[l] - kNumFunc) << 16; }
        else if (f.state == 3) {
            f.state = 4;
            const int32_t child = f.first_is_left ? r : l;
            frames.push_back(TreeFrame{child, 0, 1, 0, 0});
            continue;
        } else {
            --spills;
            code = K_POP | ((uint32_t)spills << kLevelShift) | (f.first_is_left ? rev : tok);
        }
        out[n_out].code = code;
        out[n_out].ab = ab;
        ++n_out;
        frames.pop_back();
    }
    return true;
}

inline bool prefix_to_postfix(const int32_t *tree, int32_t len, int32_t n_symbols,
                              std::vector<uint16_t> &out, int32_t &max_sp, std::string &err)
{
    std::vector<PrefixFrame> stack;
    const char *e = "";
    const bool ok = prefix_to_postfix(tree, len, n_symbols, out, max_sp, e, stack);
    if (!ok) err = e;
    return ok;
}

}  // namespace gsgp
