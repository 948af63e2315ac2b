"""ORACLE (test infrastructure only) -- Kuhn-Munkres assignment, restated.

The reference's person grouping calls ``Munkres().compute(cost)`` from the PyPI
package ``munkres`` (reference call sites: rtpe/third_party/group.py:14, :19-23,
:80).  That package is NOT vendored under /root/reference, is not installed in
this image and cannot be installed (no network); the reference pins no version
(README.md:10 defers to HigherHRNet's unpinned requirements; the contemporary
release in May 2020 was munkres 1.1.2).  This file restates the published
algorithm of munkres 1.1.x from its documented 6-step structure, keeping every
order-defining detail that decides WHICH optimal assignment is returned on ties:

* the matrix is squared with zeros (extra rows are all-zero rows);
* step 1 subtracts the row minimum (float64);
* step 2 stars the first zero of each row whose row and column are still free;
* step 4's zero search starts at the (row, col) of the previous hit, walks rows
  cyclically, and inside a row walks columns cyclically WITHOUT early exit, so
  the LAST uncovered zero (in cyclic order) of the first row that has one wins;
* step 6 applies ``+= minval`` to covered rows and then ``-= minval`` to
  uncovered columns, both to the doubly-qualified cells (not exactly reversible
  in floating point -- reproduced as written);
* zero tests are exact ``== 0`` on float64.

``start_rule="origin"`` selects the munkres <= 1.0.x behaviour (every zero
search restarts at (0, 0)).

PARITY UNPINNED: the reference holds no test, golden vector or fixture at this
boundary, and the upstream package is absent, so this restatement is anchored
only on (a) optimal-cost agreement with scipy.optimize.linear_sum_assignment
(tests/test_oracle_munkres.py) and (b) the reference's own call sites.

Nothing outside tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.
"""
from __future__ import annotations

import sys

import numpy as np


class Munkres:
    """Drop-in for ``munkres.Munkres`` as used by group.py:19-23."""

    def __init__(self, start_rule: str = "previous"):
        assert start_rule in ("previous", "origin")
        self.start_rule = start_rule

    # -- public ------------------------------------------------------------
    def compute(self, cost_matrix):
        rows = len(cost_matrix)
        cols = len(cost_matrix[0])
        n = max(rows, cols)
        # pad_matrix(): existing rows keep their values (and dtype), missing
        # columns / rows are filled with 0.
        C = np.zeros((n, n), dtype=np.float64)
        C[:rows, :cols] = np.asarray(cost_matrix, dtype=np.float64)
        self.C = C
        self.n = n
        self.row_covered = [False] * n
        self.col_covered = [False] * n
        self.marked = [[0] * n for _ in range(n)]
        self.path = [[0, 0] for _ in range(2 * n)]
        self.Z0_r = 0
        self.Z0_c = 0

        step = 1
        steps = {1: self._step1, 2: self._step2, 3: self._step3,
                 4: self._step4, 5: self._step5, 6: self._step6}
        while step in steps:
            step = steps[step]()

        results = []
        for i in range(rows):
            for j in range(cols):
                if self.marked[i][j] == 1:
                    results.append((i, j))
        return results

    # -- the six steps -------------------------------------------------------
    def _step1(self):
        C, n = self.C, self.n
        for i in range(n):
            minval = C[i, 0]
            for j in range(1, n):
                if C[i, j] < minval:
                    minval = C[i, j]
            for j in range(n):
                C[i, j] = C[i, j] - minval
        return 2

    def _step2(self):
        C, n = self.C, self.n
        for i in range(n):
            for j in range(n):
                if C[i, j] == 0 and not self.col_covered[j] \
                        and not self.row_covered[i]:
                    self.marked[i][j] = 1
                    self.col_covered[j] = True
                    self.row_covered[i] = True
                    break
        self._clear_covers()
        return 3

    def _step3(self):
        n = self.n
        count = 0
        for i in range(n):
            for j in range(n):
                if self.marked[i][j] == 1 and not self.col_covered[j]:
                    self.col_covered[j] = True
                    count += 1
        return 7 if count >= n else 4

    def _step4(self):
        row, col = 0, 0
        while True:
            if self.start_rule == "origin":
                row, col = self._find_a_zero(0, 0)
            else:
                row, col = self._find_a_zero(row, col)
            if row < 0:
                return 6
            self.marked[row][col] = 2
            star_col = self._find_star_in_row(row)
            if star_col >= 0:
                col = star_col
                self.row_covered[row] = True
                self.col_covered[col] = False
            else:
                self.Z0_r = row
                self.Z0_c = col
                return 5

    def _step5(self):
        count = 0
        path = self.path
        path[0][0] = self.Z0_r
        path[0][1] = self.Z0_c
        while True:
            row = self._find_star_in_col(path[count][1])
            if row < 0:
                break
            count += 1
            path[count][0] = row
            path[count][1] = path[count - 1][1]
            col = self._find_prime_in_row(path[count][0])
            count += 1
            path[count][0] = path[count - 1][0]
            path[count][1] = col
        for i in range(count + 1):
            r, c = path[i]
            self.marked[r][c] = 0 if self.marked[r][c] == 1 else 1
        self._clear_covers()
        for i in range(self.n):
            for j in range(self.n):
                if self.marked[i][j] == 2:
                    self.marked[i][j] = 0
        return 3

    def _step6(self):
        C, n = self.C, self.n
        minval = sys.maxsize
        for i in range(n):
            for j in range(n):
                if not self.row_covered[i] and not self.col_covered[j]:
                    if minval > C[i, j]:
                        minval = C[i, j]
        events = 0
        for i in range(n):
            for j in range(n):
                if self.row_covered[i]:
                    C[i, j] = C[i, j] + minval
                    events += 1
                if not self.col_covered[j]:
                    C[i, j] = C[i, j] - minval
                    events += 1
                if self.row_covered[i] and not self.col_covered[j]:
                    events -= 2
        if events == 0:
            raise RuntimeError("Matrix cannot be solved!")
        return 4

    # -- helpers -----------------------------------------------------------
    def _find_a_zero(self, i0, j0):
        C, n = self.C, self.n
        row, col = -1, -1
        i = i0
        done = False
        while not done:
            j = j0
            while True:
                if C[i, j] == 0 and not self.row_covered[i] \
                        and not self.col_covered[j]:
                    row, col = i, j
                    done = True
                j = (j + 1) % n
                if j == j0:
                    break
            i = (i + 1) % n
            if i == i0:
                done = True
        return row, col

    def _find_star_in_row(self, row):
        for j in range(self.n):
            if self.marked[row][j] == 1:
                return j
        return -1

    def _find_star_in_col(self, col):
        for i in range(self.n):
            if self.marked[i][col] == 1:
                return i
        return -1

    def _find_prime_in_row(self, row):
        for j in range(self.n):
            if self.marked[row][j] == 2:
                return j
        return -1

    def _clear_covers(self):
        for i in range(self.n):
            self.row_covered[i] = False
            self.col_covered[i] = False
