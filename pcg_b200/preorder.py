"""The pre-order pass for dynamic characters.

``preorder_device`` is the product path: post-order contexts (alignments on the device, ``postorder_contexts``) go to
``pcg_b200.Preorder`` — the descent over all trees x characters as prefix-sum kernels, do_preorder.cu.  The functions
below it are a host-side mirror of the pass's primitives — the pure list scans of the reference, restated literally —
kept as a readable statement of the semantics and as a second, independent restatement for the tests.

Reference (DO = lib/core/analysis/src/Analysis/Parsimony/Dynamic/DirectOptimization):
* ``get_context``                  — ``getContext`` (lib/core/data-structures/src/Bio/Character/Encodable/Dynamic/Element.hs:140-146)
* ``derive_implied_alignment``     — ``deriveImpliedAlignment`` (DO/Internal.hs:236-286)
* ``lexically_disambiguate``       — ``lexicallyDisambiguate`` / ``disambiguateElement`` (DO/Internal.hs:186-211)
* ``preorder_implied_alignments``  — ``directOptimizationPreorder`` over a rooted tree: ``initializeRoot`` (:166-177) and
                                     ``updateFromParent`` (:216-233).  The reference drives it from the re-rooted traversal
                                     foci (``preorderFromRooting``, Bio/Graph/PhylogeneticDAG/Preorder.hs:359-520); that
                                     driver is NOT mirrored — see DESIGN.md, row (f)3: upstream offers nothing to pin it
                                     against (its golden file prints the alignment context in the implied-alignment field,
                                     its script tests are disabled where the pass fails).

A dynamic character is ``None`` (Missing) or an (n, 3) array of (median, left, right) ambiguity sets.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

GAPPING, DELETION, INSERTION, ALIGNMENT = 0, 1, 2, 3


def get_context(e) -> int:
    """(left, right) non-empty? -> Gapping / Deletion (right only) / Insertion (left only) / Alignment."""
    left, right = int(e[1]) != 0, int(e[2]) != 0
    return ALIGNMENT if (left and right) else INSERTION if left else DELETION if right else GAPPING


def derive_implied_alignment(is_left_child: bool, p_alignment: np.ndarray, p_context: np.ndarray,
                             c_context: np.ndarray, gap: int) -> np.ndarray:
    """The child's final alignment from the parent's final alignment, the parent's preliminary context and the child's
    preliminary context: one fold over the parent's alignment (same length out)."""
    gap_elem = (gap, 0, 0)                                   # gapElement
    xs = [tuple(int(v) for v in t) for t in c_context]
    ys = [tuple(int(v) for v in t) for t in p_context]
    xi = yi = 0
    ys_dropped = False
    acc = []
    for e in p_alignment:
        if xi >= len(xs):                                    # go (acc, [], _) _ = (gap : acc, [], [])
            acc.append(gap_elem)
            ys_dropped = True
            continue
        if ys_dropped or yi >= len(ys):
            raise RuntimeError("Impossible happened in 'deriveImpliedAlignment'")     # as the reference
        c = get_context(e)
        x, y = xs[xi], ys[yi]
        if c == GAPPING:
            acc.append(tuple(int(v) for v in e))
        elif c == ALIGNMENT:
            acc.append(x); xi += 1; yi += 1
        elif c == DELETION:
            if get_context(y) == DELETION and not is_left_child:
                acc.append(x); xi += 1; yi += 1
            else:
                acc.append(gap_elem); yi += 1
        else:
            if get_context(y) == INSERTION and is_left_child:
                acc.append(x); xi += 1; yi += 1
            else:
                acc.append(gap_elem); yi += 1
    dtype = np.asarray(p_alignment).dtype
    return np.array(acc, dtype=dtype).reshape(-1, 3)


def lexically_disambiguate(char: np.ndarray, a: int) -> np.ndarray:
    """Every element becomes (v, v, v) with v the single symbol at index min(a - 1, countLeadingZeros median); bv-little
    counts "leading" zeros from index 0 (pinned by the L1 script costs), so v is the LOWEST set symbol, and
    an empty median gives the gap."""
    out = np.zeros_like(char)
    for k, e in enumerate(char):
        m = int(e[0]) & ((1 << a) - 1)
        idx = min(a - 1, (m & -m).bit_length() - 1) if m else a - 1
        out[k] = (1 << idx,) * 3
    return out


def preorder_implied_alignments(children: np.ndarray, n_taxa: int, contexts: Dict[int, Optional[np.ndarray]], a: int,
                                root: Optional[int] = None):
    """Final (implied) alignments and single disambiguations of every node below ``root`` (default: the tree's root) from
    the post-order contexts, in the node numbering of ``pcg_b200.Forest`` (internal node x = n_taxa + x)."""
    gap = 1 << (a - 1)
    kids = {n_taxa + x: (int(children[x, 0]), int(children[x, 1])) for x in range(len(children))}
    if root is None:
        is_child = {c for pair in kids.values() for c in pair}
        root = next(n for n in kids if n not in is_child)
    ia = {root: contexts[root]}                                                   # initializeRoot
    single = {root: None if contexts[root] is None else lexically_disambiguate(contexts[root], a)}
    stack = [root]
    while stack:
        p = stack.pop()
        for side, c in enumerate(kids.get(p, ())):
            cac = contexts[c]
            if cac is None or ia[p] is None:                                      # isMissing: inherit the parent's
                ia[c], single[c] = ia[p], single[p]
            else:
                ia[c] = derive_implied_alignment(side == 0, ia[p], contexts[p], cac, gap)
                single[c] = lexically_disambiguate(ia[c], a)
            stack.append(c)
    return ia, single


def preorder_device(ctx, children: np.ndarray, n_taxa: int, contexts, a: int):
    """Final alignments and single disambiguations of every node, computed on the device (``pcgdo_preorder_*``).
    children: (n_trees, n_taxa-1, 2) or (n_taxa-1, 2); contexts[t][c]: node -> (n, 3) array / None, as ``postorder_contexts``
    returns them per tree and character.  Returns ia[t][c][node], single[t][c][node]."""
    from .backend import Preorder
    ch = np.asarray(children, np.int32)
    if ch.ndim == 2:
        ch, contexts = ch[None], [contexts]
    n_chars = len(contexts[0])
    pre = Preorder(ctx, ch, n_taxa, n_chars)
    try:
        pre.set_contexts(contexts)
        pre.run(a)
        ia = [[dict() for _ in range(n_chars)] for _ in range(len(ch))]
        single = [[dict() for _ in range(n_chars)] for _ in range(len(ch))]
        for t in range(len(ch)):
            for c in range(n_chars):
                for v in range(2 * n_taxa - 1):
                    ia[t][c][v], single[t][c][v] = pre.node(t, c, v)
    finally:
        pre.close()
    return ia, single


# ------------------------------------------------------------------------------------------------
# The driver over re-rooted traversal foci (trees)
# ------------------------------------------------------------------------------------------------
def rooting_contexts(ctx, tcm, children: np.ndarray, n_taxa: int, leaves):
    """Directed subtree data WITH their alignment contexts, for one tree and one character — the re-rooting pass of
    ``assignOptimalDynamicCharacterRootEdges`` (Bio/Graph/PhylogeneticDAG/DynamicCharacterRerooting.hs:230-330,
    ``contextNodeDatum``) kept with the full (median, left, right) contexts the pre-order needs instead of medians only.
    Every alignment runs on the device (``foreign_pairwise_do_batch``, one batch per round of ready data).

    Returns (root, adj, ctx_of, total_of, edges):
      adj[n]           = [c1, c2, unrooted parent] of an internal non-root node (children ascending; the sibling stands in
                         for the root)
      ctx_of[(n, q)]   = context of "n entered from q" (for a leaf, or q = n's original parent: the post-order context)
      total_of[(n, q)] = its total cost
      edges            = [(u, v, total, context of the rooting on that edge)], edge 0 = the original root edge"""
    from .pairwise import foreign_pairwise_do_batch
    from .postorder import postorder_contexts
    ch = np.asarray(children, np.int32)
    n_nodes = 2 * n_taxa - 1
    down, local, total = postorder_contexts(ctx, tcm, ch, n_taxa, leaves)
    parent = {int(c): n_taxa + x for x in range(n_taxa - 1) for c in ch[x]}
    root = next(v for v in range(n_taxa, n_nodes) if v not in parent)
    rc = [int(ch[root - n_taxa][0]), int(ch[root - n_taxa][1])]
    adj = {}
    for v in range(n_taxa, n_nodes):
        if v == root:
            continue
        c1, c2 = sorted(int(c) for c in ch[v - n_taxa])
        p = parent[v]
        if p == root:
            p = rc[1] if rc[0] == v else rc[0]
        adj[v] = [c1, c2, p]
    ctx_of, total_of = {}, {}

    def known(n, q):
        return n < n_taxa or (n, q) in ctx_of

    def get(n, q):
        return (down[n], total[n]) if n < n_taxa else (ctx_of[(n, q)], total_of[(n, q)])
    for v, (c1, c2, p) in adj.items():                       # base case: entered from the original parent
        ctx_of[(v, p)], total_of[(v, p)] = down[v], total[v]
    pending = [(n, s) for n in adj for s in (0, 1)]          # entered from child adj[n][s]: DO(datum j, datum k), rotation
    while pending:
        ready = [(n, s) for (n, s) in pending
                 if known(adj[n][(s + 1) % 3], n) and known(adj[n][(s + 2) % 3], n)]
        assert ready, "cyclic dependency among directed data"
        lhs = [get(adj[n][(s + 1) % 3], n) for n, s in ready]
        rhs = [get(adj[n][(s + 2) % 3], n) for n, s in ready]
        res = foreign_pairwise_do_batch(ctx, tcm, [x[0] for x in lhs], [x[0] for x in rhs])
        for (n, s), (cost, c), l, r in zip(ready, res, lhs, rhs):
            ctx_of[(n, adj[n][s])], total_of[(n, adj[n][s])] = c, cost + l[1] + r[1]
        done = set(ready)
        pending = [x for x in pending if x not in done]
    edges = [(rc[0], rc[1], total[root], down[root])]
    pairs = [(i, int(j)) for i in adj for j in ch[i - n_taxa]]
    lhs = [get(i, j) for i, j in pairs]
    rhs = [get(j, i) for i, j in pairs]
    res = foreign_pairwise_do_batch(ctx, tcm, [x[0] for x in lhs], [x[0] for x in rhs])
    for (i, j), (cost, c), l, r in zip(pairs, res, lhs, rhs):
        edges.append((i, j, cost + l[1] + r[1], c))
    return root, adj, ctx_of, total_of, edges


def preorder_from_rooting(ctx, tcm, children: np.ndarray, n_taxa: int, leaves, a: int, edge: Optional[int] = None):
    """The pre-order pass driven from the character's traversal focus (``preorderFromRooting``,
    Bio/Graph/PhylogeneticDAG/Preorder.hs:359-520, trees): the tree is re-rooted on its cheapest edge (``edge``: force
    one), the rooting's datum is the virtual root (``FociEdgeNode`` / ``RootContext``), every other node is decorated
    against its parent in the re-rooted orientation with the datum it has when entered from that parent, and the descent
    itself runs on the device (``pcg_b200.Preorder``).

    Returns (edge index, (u, v), total cost, children of the re-rooted tree, contexts per node, ia, single) in the node
    numbering of the original tree; the virtual root takes the id of the original root (which is no node of the re-rooted
    tree).  Parity unpinned upstream, as the rest of the pass."""
    ch = np.asarray(children, np.int32)
    root, adj, ctx_of, total_of, edges = rooting_contexts(ctx, tcm, ch, n_taxa, leaves)
    if edge is None:
        edge = min(range(len(edges)), key=lambda e: (edges[e][2], e))       # first minimal edge
    u, v, tot, root_ctx = edges[edge]
    new_children = np.zeros_like(ch)
    contexts: Dict[int, Optional[np.ndarray]] = {x: np.stack([np.asarray(leaves[x], np.uint8)] * 3, axis=1) for x in range(n_taxa)}
    contexts[root] = root_ctx
    new_children[root - n_taxa] = (u, v)
    stack = [(u, v), (v, u)]                                                # (node, entered from)
    while stack:
        n, frm = stack.pop()
        if n < n_taxa:
            continue
        contexts[n] = ctx_of[(n, frm)]
        s = adj[n].index(frm)
        j, k = adj[n][(s + 1) % 3], adj[n][(s + 2) % 3]                     # the operands of that datum, in order
        new_children[n - n_taxa] = (j, k)
        stack += [(j, n), (k, n)]
    ia, single = preorder_device(ctx, new_children, n_taxa, [contexts], a)
    return edge, (u, v), tot, new_children, contexts, ia[0][0], single[0][0]
