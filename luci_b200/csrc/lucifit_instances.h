// Pre-instantiated (n_lines, group-count bucket) pairs of the fit kernel; every model is built for each pair.
// A tie pattern with NV velocity groups and NS broadening groups runs on the smallest bucket >= max(NV, NS)
// (unused group columns are pinned by the solver).
#pragma once
#ifndef LF_INSTANCES
#define LF_INSTANCES(X) \
    X(1, 1) \
    X(2, 1) X(2, 2) \
    X(3, 1) X(3, 2) X(3, 3) \
    X(4, 1) X(4, 2) X(4, 4) \
    X(5, 1) X(5, 2) X(5, 5) \
    X(6, 1) X(6, 2) X(6, 6) \
    X(7, 1) X(7, 2) \
    X(8, 1) X(8, 2)
#endif
