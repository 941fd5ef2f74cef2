// program.hpp -- the interpreter's program format, shared by the host staging code (trees.hpp) and
// the on-device random-tree generator (plan_kernel).
//
// eval_arrays (main_cpu.go:755-775) recurses over the pointer-prefix array once per fitness case.
// The device interpreter runs an accumulator program instead: one instruction per FUNCTION node,
// with terminal operands folded into the instruction, so a tree of n nodes costs (n-1)/2 decoded
// instructions instead of n stack tokens, and the operand stack is touched only at nodes whose two
// children are both subtrees.  Instruction forms (l op r, result -> acc):
//     TT   l = term a, r = term b                 PTT  as TT, but first spill acc to the stack
//     AT   l = acc,    r = term b                 TA   l = term a, r = acc
//     SA   l = pop(),  r = acc                    T    acc = term a (a tree that is a lone terminal)
// Each node still computes exactly `left op right` on the same operand values as the reference, so
// the results are bit-identical; only the bookkeeping differs.  Operands are symbol id - 4:
// [0,nvar) = variable columns, [nvar, nvar+nconst) = constants (main_cpu.go:434-454,745-753).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GSGP_HD __host__ __device__ __forceinline__
#else
#define GSGP_HD inline
#endif

namespace gsgp {

enum { F_TT = 0, F_PTT = 1, F_AT = 2, F_TA = 3, F_SA = 4, F_T = 5 };

struct Ins { uint32_t code;  /* form << 2 | op */  uint32_t ab;  /* a | b << 16 */ };

constexpr uint16_t kAbsAcc = 0xFFFF, kAbsSpill = 0xFFFE;
constexpr int kMaxSymbols = 0xFFF0;                    // symbol ids must stay below the markers
constexpr int kMaxStackLevels = 32;                    // spill levels the interpreter can be given

// postfix (left subtree, right subtree, op; u16 symbol ids) -> accumulator program.
// abs_stack: scratch of abs_cap entries (>= operand-stack depth of the postfix program).
// Returns false on a malformed program or when abs_cap is too small.
GSGP_HD bool compile_postfix(const uint16_t *post, int len, Ins *out, int &n_out, int &stack_need,
                             uint16_t *abs_stack, int abs_cap)
{
    int sp = 0, acc_at = -1, spills = 0;
    n_out = 0; stack_need = 0;
    for (int i = 0; i < len; ++i) {
        const uint16_t tok = post[i];
        if (tok >= 4) {
            if (sp >= abs_cap) return false;
            abs_stack[sp++] = tok;
            continue;
        }
        if (sp < 2) return false;
        const uint16_t r = abs_stack[sp - 1], l = abs_stack[sp - 2];
        sp -= 2;
        uint32_t form, a = 0, b = 0;
        if (l < kAbsSpill && r < kAbsSpill) {
            form = F_TT;
            if (acc_at >= 0) {                          // a pending left value lives in acc: spill it
                form = F_PTT;
                abs_stack[acc_at] = kAbsSpill;
                if (++spills > stack_need) stack_need = spills;
            }
            a = l - 4u; b = r - 4u;
        } else if (l == kAbsAcc && r < kAbsSpill) { form = F_AT; b = r - 4u; }
        else if (l < kAbsSpill && r == kAbsAcc) { form = F_TA; a = l - 4u; }
        else if (l == kAbsSpill && r == kAbsAcc) { form = F_SA; --spills; }
        else return false;
        out[n_out].code = (form << 2) | tok;
        out[n_out].ab = a | (b << 16);
        ++n_out;
        abs_stack[sp] = kAbsAcc; acc_at = sp; ++sp;
    }
    if (sp != 1) return false;
    if (n_out == 0) {                                   // lone terminal (zero_depth trees)
        out[0].code = (uint32_t)F_T << 2;
        out[0].ab = abs_stack[0] - 4u;
        n_out = 1;
    }
    return true;
}

}  // namespace gsgp
