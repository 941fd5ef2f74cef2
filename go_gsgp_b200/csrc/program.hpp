// program.hpp -- the interpreter's program format, shared by the host staging code (trees.hpp) and
// the on-device random-tree generator (plan_draw_kernel).
//
// eval_arrays (main_cpu.go:755-775) recurses over the pointer-prefix array once per fitness case.
// The device interpreter runs an accumulator program instead: one instruction per FUNCTION node,
// with terminal operands folded into the instruction, so a tree of n nodes costs (n-1)/2 decoded
// instructions instead of n stack tokens, and a spill stack is touched only at nodes whose two
// children are both subtrees.  An instruction is  [push] [load a] ; t = b or pop ; acc = acc (aop) t :
//   K_LOADA          acc = term a          (both children are terminals)
//   K_PUSH           spill acc to level L first (a left sibling's value is pending; implies K_LOADA)
//   K_POP            t = spill level L     (both children subtrees: the left value was spilled)
//   otherwise        t = term b            (one child is a terminal, the other one is in acc)
//   aop    ADD SUB MUL DIV  acc = acc op t      RSUB RDIV  acc = t op acc      MOV  acc = t (lone terminal)
// Spill levels are assigned when the program is built.  Each node still computes exactly
// `left op right` on the same operand values as the reference (RSUB/RDIV keep the operand order when
// the left operand is the terminal), so results are bit-identical; only the bookkeeping differs.
// Operands are symbol id - 4: [0,nvar) = variable columns, [nvar, nvar+nconst) = constants
// (main_cpu.go:434-454,745-753).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GSGP_HD __host__ __device__ __forceinline__
#else
#define GSGP_HD inline
#endif

namespace gsgp {

enum { A_ADD = 0, A_SUB = 1, A_MUL = 2, A_DIV = 3, A_RSUB = 4, A_RDIV = 5, A_MOV = 6 };
enum { K_LOADA = 8, K_PUSH = 16, K_POP = 32 };

constexpr int kLevelShift = 11;                         // spill level pre-scaled to its byte offset: levels are 256 cases x 8 B apart
struct Ins { uint32_t code;  /* level << kLevelShift | K_* flags | aop */  uint32_t ab;  /* a | b << 16 */ };

// per tree-slot descriptor word: instruction count | spill levels needed << 24
GSGP_HD int32_t prog_desc(int n_ins, int stack_need) { return n_ins | (stack_need << 24); }
GSGP_HD int prog_desc_len(int32_t d) { return d & 0xFFFFFF; }
GSGP_HD int prog_desc_need(int32_t d) { return (d >> 24) & 0xFF; }

constexpr int kMaxSymbols = 0xFFF0;
constexpr int kMaxStackLevels = 32;                    // spill levels the interpreter can be given

// postfix (left subtree, right subtree, op; u16 symbol ids) -> accumulator program.
//
// Where both children of a node are subtrees, the one that needs more spill levels is evaluated
// first (Sethi-Ullman order), which keeps almost every random tree within two spill levels; the
// node still computes `left op right` on the same two values, so results do not change.
//   scratch : len entries; per token: subtree size | spill need << 16
//   frames  : frame_cap entries (>= tree depth + 1)
// Returns false on a malformed program or when frame_cap is too small.
struct CompileFrame { uint16_t node; uint8_t state, live, first_is_left; };

// PostA / ScratchA: anything indexable with [] (plain pointers on the host; bank-interleaved shared-memory views in
// plan_draw_kernel, where a thread-serial walk over global memory would be latency-bound).
template <typename PostA, typename ScratchA>
GSGP_HD bool compile_postfix(PostA post, int len, Ins *out, int &n_out, int &stack_need,
                             ScratchA scratch, CompileFrame *frames, int frame_cap)
{
    n_out = 0; stack_need = 0;
    if (len <= 0) return false;
    // pass 1: subtree sizes and spill needs, bottom-up.  The operand stack of the postfix walk is
    // threaded through scratch itself: a finished subtree's root position is all that is needed.
    {
        int depth = 0;                                  // operand-stack depth (validation only)
        for (int i = 0; i < len; ++i) {
            if (post[i] >= 4) { scratch[i] = 1u; ++depth; continue; }
            if (depth < 2) return false;
            const int r = i - 1;
            const uint32_t sr = scratch[r] & 0xFFFFu, nr = scratch[r] >> 16;
            const int l = r - (int)sr;
            if (l < 0) return false;
            const uint32_t sl = scratch[l] & 0xFFFFu, nl = scratch[l] >> 16;
            const bool lt = post[l] >= 4, rt = post[r] >= 4;
            uint32_t need;
            if (!lt && !rt) need = (nl == nr) ? nl + 1u : (nl > nr ? nl : nr);
            else need = lt ? nr : nl;                   // a terminal child needs nothing (nr = 0 for terminals)
            if (1u + sl + sr > 0xFFFFu) return false;
            scratch[i] = (1u + sl + sr) | (need << 16);
            --depth;
        }
        if (depth != 1 || (int)(scratch[len - 1] & 0xFFFFu) != len) return false;
    }
    if (post[len - 1] >= 4) {                           // lone terminal (zero_depth trees)
        out[0].code = (uint32_t)A_MOV;
        out[0].ab = (uint32_t)(post[0] - 4u) << 16;
        n_out = 1;
        return true;
    }
    // pass 2: emit in evaluation order
    int nf = 0, spills = 0;
    if (frame_cap < 1) return false;
    frames[nf++] = CompileFrame{(uint16_t)(len - 1), 0, 0, 0};
    while (nf > 0) {
        CompileFrame &f = frames[nf - 1];
        const int i = f.node, r = i - 1, l = r - (int)(scratch[r] & 0xFFFFu);
        const uint32_t tok = post[i];
        // tok: 0 + 1 - 2 * 3 / (main_cpu.go:434-437) = A_ADD..A_DIV; reversed forms when acc is the RIGHT operand
        const uint32_t rev = (tok == 1) ? (uint32_t)A_RSUB : (tok == 3) ? (uint32_t)A_RDIV : tok;
        const bool lt = post[l] >= 4, rt = post[r] >= 4;
        uint32_t code, ab = 0;
        if (f.state == 0) {
            if (lt && rt) {
                code = K_LOADA | tok;
                if (f.live) {                           // a pending sibling value lives in acc: spill it
                    code |= K_PUSH | ((uint32_t)spills << kLevelShift);
                    if (++spills > stack_need) stack_need = spills;
                }
                ab = (uint32_t)(post[l] - 4u) | ((uint32_t)(post[r] - 4u) << 16);
            } else {
                int child;
                if (!lt && rt) { f.state = 1; child = l; }
                else if (lt && !rt) { f.state = 2; child = r; }
                else {
                    f.first_is_left = (scratch[r] >> 16) > (scratch[l] >> 16) ? 0 : 1;
                    f.state = 3;
                    child = f.first_is_left ? l : r;
                }
                if (nf >= frame_cap) return false;
                const uint8_t live = f.live;
                frames[nf++] = CompileFrame{(uint16_t)child, 0, live, 0};
                continue;
            }
        } else if (f.state == 1) { code = tok; ab = (uint32_t)(post[r] - 4u) << 16; }
        else if (f.state == 2) { code = rev; ab = (uint32_t)(post[l] - 4u) << 16; }
        else if (f.state == 3) {
            f.state = 4;
            if (nf >= frame_cap) return false;
            const int child = f.first_is_left ? r : l;
            frames[nf++] = CompileFrame{(uint16_t)child, 0, 1, 0};
            continue;
        } else {                                        // popped value = the child evaluated first
            --spills;
            code = K_POP | ((uint32_t)spills << kLevelShift) | (f.first_is_left ? rev : tok);
        }
        out[n_out].code = code;
        out[n_out].ab = ab;
        ++n_out;
        --nf;
    }
    return true;
}

}  // namespace gsgp
