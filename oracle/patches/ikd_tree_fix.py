#!/usr/bin/env python
"""Repairs the brace damage of the reference's src/ikd-Tree/ikd_Tree.cpp so that it compiles (TEST INFRASTRUCTURE).

The fork's Korean-comment annotation pass broke the block structure of ikd_Tree.cpp in five places and duplicated one block (SURVEY.md "Read
this first" item 3: `g++ -c` fails at :675, :1250, :1884, ...).  Each edit below restores upstream ikd-Tree's structure
at that place; nothing else of the file is touched.  The edits are addressed by line number of the pinned file (its
SHA-256 is checked), so this script carries no reference text beyond the lines it has to ADD:

  :301-305 multi_thread_rebuild  the "wait until no search is running" block (:292-297) appears a second time and
                               re-locks search_flag_mutex, a default (non-recursive) mutex the thread already holds: the
                               rebuild thread dead-locks on the first sub-tree rebuild of >= 1500 points
                               (Multi_Thread_Rebuild_Point_Num) and every later Add_Points / Nearest_Search blocks behind
                               it.  Upstream has the block once                              -> the five duplicated lines dropped
  :675   Add_Points            the `if (Rebuild_Ptr == nullptr || ...) {...} else {...}` pair inside `if (downsample_switch) {`
                               is not closed before the `else` of downsample_switch          -> one `}` added
  :1246-1248 Add_by_range      `}` / `return;` / `}` : the early return of the "box contains the whole node" case fell
                               outside its `if`, and the function was closed there           -> return moved inside, `}` dropped
  :1880  Push_Down             the `if (root->need_push_down_to_right ...) {` block is not closed before `return;`  -> one `}` added
  :1918  Update (both sons)    the test that selects the cheap AABB merge lost its `if (...) {` line (the two single-son
                               branches :1982 / :2045 still have theirs; this is their two-son form, as upstream)
  :1997  Update (left son)     `else {` lost its inner `if (!left_son_ptr->tree_deleted) {` (the right-son branch :2062 kept it)

usage: ikd_tree_fix.py <reference ikd_Tree.cpp> <output .cpp>      (the output goes to a temporary build directory; the
patched source is never stored in this repository -- only oracle/_ref/libref_ikd.so is kept, git-ignored)
"""
import hashlib
import sys

PINNED_SHA256 = "5be4be2d68ce75bccbb19c8aceb3cfa8103213c266d355ea52e6e7a1480ed5bd"

EDITS = [  # (1-based line of the pinned file, operation, text)
    (301, "delete", None), (302, "delete", None), (303, "delete", None), (304, "delete", None), (305, "delete", None),
    (675, "before", "        }"),
    (1246, "replace", "        return;"),
    (1247, "replace", "    }"),
    (1248, "delete", None),
    (1880, "after", "}"),
    (1918, "before", "    if (root->tree_deleted || (!left_son_ptr->tree_deleted && !right_son_ptr->tree_deleted && "
                     "!root->point_deleted)) {"),
    (1997, "after", "        if (!left_son_ptr->tree_deleted) {"),
]


def main(src, dst):
    raw = open(src, "rb").read()
    got = hashlib.sha256(raw).hexdigest()
    if got != PINNED_SHA256:
        sys.exit(f"ikd_tree_fix: {src} is not the pinned file (sha256 {got}); the line-addressed edits do not apply")
    lines = raw.decode("utf-8").split("\n")
    for ln, op, txt in sorted(EDITS, key=lambda e: -e[0]):  # bottom-up: earlier line numbers stay valid
        i = ln - 1
        if op == "after":
            lines.insert(i + 1, txt)
        elif op == "before":
            lines.insert(i, txt)
        elif op == "replace":
            lines[i] = txt
        elif op == "delete":
            del lines[i]
    open(dst, "w", encoding="utf-8").write("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
