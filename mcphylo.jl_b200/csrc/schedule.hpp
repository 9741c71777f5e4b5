// schedule.hpp — host-side compiler from a tree topology to the flat "walk program" the
// device executes.  Pure C++ (no CUDA) so it can be exercised without a GPU
// (mcp_schedule_dump, tests/test_schedule_emulator.py).
//
// Every alignment column is owned by ONE thread for the whole evaluation: the thread walks all
// nodes itself, so the program is a straight list of ops with no inter-thread dependencies.
// The reference walks the same nodes one dense array at a time
// (/root/reference/src/Likelihood/LikelihoodCalculator_Node.jl:92-111 post-order,
//  :25-74 reverse post-order); the traversal ORDER here is chosen for locality, the arithmetic
// per node is the reference's.
//
// Device tree: strictly binary.  A node with k >= 3 children (the reference's tests have a
// trifurcating root) is folded left into k-2 virtual nodes on identity branches, which keeps
// the reference's product order ((D1*D2)*D3..); a node with one child gets a virtual all-ones
// leaf on an identity branch.  Virtual branches have P = I, dP = 0 and no gradient output.
//
// POST program, one op per internal device node, executed in an order where the child with the
// larger subtree comes first.  Then the second child (if internal) is always the op executed
// immediately before (operand kind REG, still in registers) and only the first child may have
// to be re-read (kind MEM).  In gradient mode every non-root result is stored (slot = op
// index) because the pre pass needs it; in logL-only mode only results that will be re-read
// are stored, into a LIFO of depth <= log2(N)+1.
//
// PRE program (gradient mode), one op per internal device node = one family {mother, a, b}:
//   Da = P_a L_a, Db = P_b L_b, qa = pre_m * Db, qb = pre_m * Da, den = sum pre_m*Da*Db,
//   grad_a += qa.(dP_a L_a)/den, grad_b likewise, pre_a = P_a^T qa, pre_b = P_b^T qb.
// Depth-first, descending into the child with the SMALLER subtree first (its pre vector stays
// in registers: KEEP); the other internal child is pushed on a LIFO (PUSH) of depth
// <= log2(N)+1 and popped when the walk comes back.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

namespace mcp {

enum : int { OPK_LEAF = 0, OPK_REG = 1, OPK_MEM = 2, OPK_CHERRY = 3 };   // operand kinds (CHERRY: pre program only)
enum : int { PREM_ROOT = 0, PREM_REG = 1, PREM_STACK = 2 };     // where pre[mother] comes from
enum : int { OUT_NONE = 0, OUT_KEEP = 1, OUT_PUSH = 2 };        // what happens to pre[child]

// CHERRY (pre program, child a only): the child is an internal node whose two children are both real
// leaves x, y.  Its post result L = rescale(P_x[:, code_x] * P_y[:, code_y]) is two table look-ups and K
// multiplies, so the gradient pass RECOMPUTES it from the codes instead of re-reading a stored copy, and the
// post pass does not store it: in a random binary tree a third of the internal nodes are cherries, i.e. a
// third of the stored partials and of their re-reads (64 of ~190 bytes per column and node at K = 4).
//   a_src = row_x | row_y << 16   (alignment rows)     a_dst = br_x | br_y << 16   (branch-table rows)
//   a_br stays the cherry's own branch.  Same arithmetic in the same order as the post op, hence the same
//   bits.  Only emitted when allow_cherry is set (depth-first walk kernels for K <= 6), for a cherry that is
//   the kept child of its family and whose post result the post pass itself does not re-read.
//
// Both op types are 8 x int32 with a common head so that the kernel's staging code can treat
// them alike:
//   word 0..3  a_src, a_br, b_src, b_br   child operands: LEAF -> alignment row (-1 = all ones),
//                                          MEM -> post slot of the child, REG -> unused;
//                                          *_br = device node id of the child (branch-table row)
//   word 4     post: dst slot (if STORE)   pre: m_src (LIFO slot when the mother is on the stack)
//   word 5     flags: bits 0-1 a kind, 2-3 b kind (OPK_*), 4 POST_STORE, 5 POST_ROOT,
//              pre only: bits 8-9 where pre[mother] comes from (PREM_*), 10-11 a out, 12-13 b out (OUT_*)
//   word 6     post: device node id (debug) pre: a_dst (LIFO slot if a is pushed)
//   word 7     pre: b_dst
struct PostOp {
    int32_t a_src, a_br, b_src, b_br;
    int32_t dst;
    int32_t flags;
    int32_t node;
    int32_t pad;
};
struct PreOp {
    int32_t a_src, a_br, b_src, b_br;
    int32_t m_src;
    int32_t flags;
    int32_t a_dst, b_dst;
};
static_assert(sizeof(PostOp) == 32 && sizeof(PreOp) == 32, "ops are two 16-byte words");

constexpr int POST_STORE = 1 << 4, POST_ROOT = 1 << 5;

struct Schedule {
    std::vector<PostOp> post;
    std::vector<PreOp> pre;
    int n_dnodes = 0;   // device nodes = NN + virtual ones; also the number of branch-table rows
    int n_slots = 0;    // post result slots needed per column
    int n_stack = 0;    // pre LIFO depth needed per column
    int n_real_branches = 0;  // NN-1
    int n_cherries = 0;       // pre-program children recomputed from their two leaves instead of re-read
    // Depth-first gradient program only: the post slots of the stored child partials the pre program reads,
    // in the order it reads them (per op: child a if MEM, then child b if MEM) -- the fetch list of the walk
    // kernel's operand ring, which requests them a few entries ahead of their use.  Empty when a slot index
    // does not fit 16 bits.
    std::vector<uint16_t> pre_fetch;
    // level-ordered variant only: ops [levels[l], levels[l+1]) are mutually independent
    std::vector<int32_t> post_levels, pre_levels;
};

// leaf_row[num-1] = alignment row for leaves (must be >= 0 for every childless node).
// Returns "" on success, else an error message.
//
// by_levels = true emits the LEVEL-ORDERED variant for the small-tree latency kernel: the same op
// semantics, but ops sorted by height (post) / depth (pre) so that all ops of a level can run in
// parallel on different warps; every operand goes through a slot (no REG / KEEP), post result of
// internal node i lives in post slot i, its pre vector in pre ("LIFO") slot i.
inline std::string build_schedule(int NN, const int32_t* postorder_num, const int32_t* parent_num,
                                  const int32_t* leaf_row, bool want_grad, Schedule& out, bool by_levels = false,
                                  bool allow_cherry = false) {
    if (NN < 2) return "tree must have at least two nodes";
    if (postorder_num[NN - 1] != NN) return "root must come last in post-order and carry num == NN";
    std::vector<int> pos(NN, -1);
    for (int i = 0; i < NN; ++i) {
        int num = postorder_num[i];
        if (num < 1 || num > NN) return "postorder_num out of range";
        if (pos[num - 1] >= 0) return "postorder_num has duplicates";
        pos[num - 1] = i;
    }
    if (parent_num[NN - 1] != 0) return "root must have parent 0";
    std::vector<int> nchild(NN, 0), cstart(NN + 1, 0);
    for (int n = 0; n < NN - 1; ++n) {
        int p = parent_num[n];
        if (p < 1 || p > NN) return "parent_num out of range";
        if (pos[p - 1] <= pos[n]) return "a mother must come after her children in post-order";
        nchild[p - 1]++;
    }
    for (int n = 0; n < NN; ++n) cstart[n + 1] = cstart[n] + nchild[n];
    std::vector<int> clist(NN > 1 ? NN - 1 : 1), cfill(NN, 0);
    for (int i = 0; i < NN; ++i) {  // stored child order == order of appearance in post-order
        int n = postorder_num[i] - 1;
        int p = parent_num[n];
        if (p > 0) clist[cstart[p - 1] + cfill[p - 1]++] = n;
    }

    // ---- binarise -------------------------------------------------------------------------
    struct DNode { int left = -1, right = -1, row = -1, isize = 0; };
    std::vector<DNode> dn(NN);
    dn.reserve(2 * NN);
    for (int n = 0; n < NN; ++n) {
        int k = nchild[n];
        const int* ch = clist.data() + cstart[n];
        if (k == 0) {
            if (leaf_row[n] < 0) return "leaf node " + std::to_string(n + 1) + " has no alignment row";
            dn[n].row = leaf_row[n];
        } else if (k == 1) {
            DNode one;  // virtual all-ones leaf on an identity branch
            one.row = -1;
            dn.push_back(one);
            dn[n].left = ch[0];
            dn[n].right = (int)dn.size() - 1;
        } else {
            int acc = ch[0];
            for (int i = 1; i < k - 1; ++i) {
                DNode v;
                v.left = acc;
                v.right = ch[i];
                dn.push_back(v);
                acc = (int)dn.size() - 1;
            }
            dn[n].left = acc;
            dn[n].right = ch[k - 1];
        }
    }
    const int ND = (int)dn.size();
    const int root = NN - 1;
    auto is_leaf = [&](int d) { return dn[d].left < 0; };

    // ---- subtree sizes (internal device nodes), iterative post-order ----------------------
    std::vector<int> order;  // generic post-order of internal nodes, for isize
    order.reserve(ND);
    {
        std::vector<std::pair<int, int>> st;
        st.push_back({root, 0});
        while (!st.empty()) {
            auto& [d, state] = st.back();
            if (is_leaf(d)) { st.pop_back(); continue; }
            if (state == 0) { state = 1; int c = dn[d].left; st.push_back({c, 0}); }
            else if (state == 1) { state = 2; int c = dn[d].right; st.push_back({c, 0}); }
            else { dn[d].isize = 1 + dn[dn[d].left].isize + dn[dn[d].right].isize; order.push_back(d); st.pop_back(); }
        }
    }

    out = Schedule();
    out.n_dnodes = ND;
    out.n_real_branches = NN - 1;
    const int n_int = (int)order.size();
    out.post.reserve(n_int);
    std::vector<int> slot_of(ND, -1);  // where a node's post result lives

    if (by_levels) {
        std::vector<int> height(ND, 0), depth(ND, 0);
        for (int d : order) height[d] = 1 + std::max(height[dn[d].left], height[dn[d].right]);
        std::vector<int> byh(order);
        std::stable_sort(byh.begin(), byh.end(), [&](int x, int y) { return height[x] < height[y]; });
        for (int i = 0; i < n_int; ++i) slot_of[byh[i]] = i;
        for (int i = 0; i < n_int; ++i) {
            const int d = byh[i];
            PostOp op{};
            op.node = d;
            auto operand = [&](int c, int32_t& src, int32_t& br) -> int {
                br = c;
                if (is_leaf(c)) { src = dn[c].row; return OPK_LEAF; }
                src = slot_of[c];
                return OPK_MEM;
            };
            int ka = operand(dn[d].left, op.a_src, op.a_br);
            int kb = operand(dn[d].right, op.b_src, op.b_br);
            op.flags = ka | (kb << 2);
            if (d == root) op.flags |= POST_ROOT;
            else { op.flags |= POST_STORE; op.dst = i; }
            if (i == 0 || height[d] != height[byh[i - 1]]) out.post_levels.push_back(i);
            out.post.push_back(op);
        }
        out.post_levels.push_back(n_int);
        out.n_slots = n_int;
        if (want_grad) {
            for (auto it = order.rbegin(); it != order.rend(); ++it) {   // mothers before children
                depth[dn[*it].left] = depth[*it] + 1;
                depth[dn[*it].right] = depth[*it] + 1;
            }
            std::vector<int> byd(order.rbegin(), order.rend());
            std::stable_sort(byd.begin(), byd.end(), [&](int x, int y) { return depth[x] < depth[y]; });
            for (int i = 0; i < n_int; ++i) {
                const int m = byd[i];
                PreOp op{};
                const int a = dn[m].left, b = dn[m].right;
                const bool ai = !is_leaf(a), bi = !is_leaf(b);
                op.a_br = a; op.b_br = b;
                op.a_src = ai ? slot_of[a] : dn[a].row;
                op.b_src = bi ? slot_of[b] : dn[b].row;
                op.m_src = slot_of[m];
                op.a_dst = ai ? slot_of[a] : 0;
                op.b_dst = bi ? slot_of[b] : 0;
                op.flags = (ai ? OPK_MEM : OPK_LEAF) | ((bi ? OPK_MEM : OPK_LEAF) << 2) |
                           ((m == root ? PREM_ROOT : PREM_STACK) << 8) |
                           ((ai ? OUT_PUSH : OUT_NONE) << 10) | ((bi ? OUT_PUSH : OUT_NONE) << 12);
                if (i == 0 || depth[m] != depth[byd[i - 1]]) out.pre_levels.push_back(i);
                out.pre.push_back(op);
            }
            out.pre_levels.push_back(n_int);
            out.n_stack = n_int;
        }
        return "";
    }

    std::vector<char> post_reread(ND, 0);   // the post pass itself re-reads this node's stored result
    // ---- POST program: larger subtree first --------------------------------------------------
    {
        std::vector<std::pair<int, int>> st;
        st.push_back({root, 0});
        int last_emitted = -1, depth = 0, max_depth = 0;
        while (!st.empty()) {
            auto [d, state] = st.back();
            if (is_leaf(d)) { st.pop_back(); continue; }
            int l = dn[d].left, r = dn[d].right;
            bool left_first = dn[l].isize >= dn[r].isize;
            int first = left_first ? l : r, second = left_first ? r : l;
            if (state == 0) { st.back().second = 1; st.push_back({first, 0}); continue; }
            if (state == 1) {
                st.back().second = 2;
                // logL-only: the first child's result must be parked if the second child's
                // subtree will overwrite the registers
                if (!want_grad && !is_leaf(first) && !is_leaf(second)) {
                    PostOp& fop = out.post[slot_of[first]];  // slot_of holds the op index for now
                    fop.flags |= POST_STORE;
                    fop.dst = depth++;
                    if (depth > max_depth) max_depth = depth;
                }
                st.push_back({second, 0});
                continue;
            }
            st.pop_back();
            PostOp op{};
            op.node = d;
            auto operand = [&](int c, int32_t& src, int32_t& br) -> int {
                br = c;
                if (is_leaf(c)) { src = dn[c].row; return OPK_LEAF; }
                if (c == last_emitted) { src = 0; return OPK_REG; }
                src = want_grad ? slot_of[c] : out.post[slot_of[c]].dst;
                post_reread[c] = 1;
                return OPK_MEM;
            };
            int ka = operand(l, op.a_src, op.a_br);
            int kb = operand(r, op.b_src, op.b_br);
            // Canonical operand order (the product Da * Db is commutative, bit for bit): the only
            // kind pairs that occur are (LEAF, LEAF), (REG, LEAF) and (MEM, REG) -- the walk
            // kernel branches three ways on a's kind alone.
            if ((ka == OPK_LEAF && kb != OPK_LEAF) || (ka == OPK_REG && kb == OPK_MEM)) {
                std::swap(ka, kb);
                std::swap(op.a_src, op.b_src);
                std::swap(op.a_br, op.b_br);
            }
            op.flags = ka | (kb << 2);
            if (!want_grad && (ka == OPK_MEM || kb == OPK_MEM)) depth--;  // LIFO pop
            if (d == root) op.flags |= POST_ROOT;
            int idx = (int)out.post.size();
            if (want_grad) {
                if (d != root) { op.flags |= POST_STORE; op.dst = idx; }
            }
            slot_of[d] = idx;
            out.post.push_back(op);
            last_emitted = d;
        }
        out.n_slots = want_grad ? (int)out.post.size() : max_depth;
        if (out.n_slots < 1) out.n_slots = 1;
    }

    // ---- PRE program: smaller subtree first, the other child parked on a LIFO ----------------
    if (want_grad) {
        out.pre.reserve(n_int);
        struct Pending { int node, slot; };
        std::vector<Pending> pending;
        int level = 0, max_level = 0;
        int cur = root, cur_kind = PREM_ROOT, cur_src = 0;
        while (true) {
            PreOp op{};
            int a = dn[cur].left, b = dn[cur].right;
            op.m_src = cur_src;
            op.a_br = a; op.b_br = b;
            bool ai = !is_leaf(a), bi = !is_leaf(b);
            op.a_src = ai ? slot_of[a] : dn[a].row;
            op.b_src = bi ? slot_of[b] : dn[b].row;
            int a_out = OUT_NONE, b_out = OUT_NONE;
            int next = -1;
            if (cur_kind == PREM_STACK) level--;  // popped; this op may reuse the slot
            if (ai && bi) {
                bool a_small = dn[a].isize <= dn[b].isize;
                int keep = a_small ? a : b, push = a_small ? b : a;
                (a_small ? a_out : b_out) = OUT_KEEP;
                (a_small ? b_out : a_out) = OUT_PUSH;
                (a_small ? op.b_dst : op.a_dst) = level;
                pending.push_back({push, level});
                level++;
                if (level > max_level) max_level = level;
                next = keep;
            } else if (ai) { a_out = OUT_KEEP; next = a; }
            else if (bi) { b_out = OUT_KEEP; next = b; }
            // Canonical child order (the family form is symmetric in a and b): a is the child whose
            // pre vector stays in registers (KEEP), b the pushed one or a leaf.  The kinds that occur
            // are (LEAF, LEAF), (MEM/KEEP, LEAF) and (MEM/KEEP, MEM/PUSH).
            if (b_out == OUT_KEEP) {
                std::swap(ai, bi);
                std::swap(a_out, b_out);
                std::swap(op.a_src, op.b_src);
                std::swap(op.a_br, op.b_br);
                std::swap(op.a_dst, op.b_dst);
            }
            int a_kind = ai ? OPK_MEM : OPK_LEAF;
            if (allow_cherry && ai && a_out == OUT_KEEP && ND < 65536) {
                const int c = op.a_br, x = dn[c].left, y = dn[c].right;
                if (is_leaf(x) && is_leaf(y) && dn[x].row >= 0 && dn[y].row >= 0 && dn[x].row < 65536 && dn[y].row < 65536 &&
                    !post_reread[c]) {
                    a_kind = OPK_CHERRY;
                    op.a_src = dn[x].row | (dn[y].row << 16);
                    op.a_dst = x | (y << 16);
                    out.post[slot_of[c]].flags &= ~POST_STORE;      // nobody reads the stored copy any more
                    ++out.n_cherries;
                }
            }
            op.flags = a_kind | ((bi ? OPK_MEM : OPK_LEAF) << 2) | (cur_kind << 8) | (a_out << 10) | (b_out << 12);
            out.pre.push_back(op);
            if (next >= 0) { cur = next; cur_kind = PREM_REG; cur_src = 0; }
            else if (!pending.empty()) {
                Pending p = pending.back(); pending.pop_back();
                cur = p.node; cur_kind = PREM_STACK; cur_src = p.slot;
            } else break;
        }
        out.n_stack = max_level < 1 ? 1 : max_level;
        if (out.n_slots < 65535) {
            for (const PreOp& op : out.pre) {
                if ((op.flags & 3) == OPK_MEM) out.pre_fetch.push_back((uint16_t)op.a_src);
                if (((op.flags >> 2) & 3) == OPK_MEM) out.pre_fetch.push_back((uint16_t)op.b_src);
            }
        }
    }
    return "";
}

}  // namespace mcp
