#!/usr/bin/env python3
"""Golden pair for the bins -> regions collapse (K0), taken verbatim from the reference's bundled tutorial data:
docs/example_data/5nodes_10regions_10000reads_test_d_mat.csv (200 cells x 1000 bins) and the segmented counts /
region sizes the reference pipeline produced from it (..._segmented_counts.txt, test_segmented_region_sizes.txt).
Writes tests/golden/tutorial_bins.npz.  TEST INFRASTRUCTURE; needs /root/reference, the output is committed."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EX = os.path.join(os.environ.get("SCICONE_REF", "/root/reference"), "docs", "example_data")
bins = np.loadtxt(os.path.join(EX, "5nodes_10regions_10000reads_test_d_mat.csv"), delimiter=",")
regions = np.loadtxt(os.path.join(EX, "5nodes_10regions_10000reads_test_d_mat_segmented_counts.txt"), delimiter=",")
sizes = np.loadtxt(os.path.join(EX, "test_segmented_region_sizes.txt")).astype(np.int32)
assert np.array_equal(bins, np.round(bins)) and bins.max() < 65536
np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "tutorial_bins.npz"), bins=bins.astype(np.uint16),
                    regions=regions, region_sizes=sizes)
print(bins.shape, regions.shape, sizes)

# cluster condensation golden: the reference's own condensed matrix / sizes for the tutorial's cluster assignments
assign = np.loadtxt(os.path.join(EX, "5nodes_10regions_10000reads_test_d_mat_segmented_counts_cluster_assignments.txt")).astype(np.int32)
condensed = np.loadtxt(os.path.join(EX, "5nodes_10regions_10000reads_test_d_mat_segmented_counts_condensed.txt"), delimiter=",")
csizes = np.loadtxt(os.path.join(EX, "5nodes_10regions_10000reads_test_d_mat_segmented_counts_cluster_sizes.txt")).astype(np.int32)
np.savez_compressed(os.path.join(HERE, "..", "tests", "golden", "tutorial_clusters.npz"), regions=regions, assignments=assign,
                    condensed=condensed, cluster_sizes=csizes)
print(condensed.shape, csizes)
