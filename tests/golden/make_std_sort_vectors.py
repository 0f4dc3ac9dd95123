"""Known-answer vectors for the oracle's restatement of libstdc++ std::sort (oracle/vracer_oracle.py, std_sort): keys with
many ties sorted by g++'s own std::sort (introsort path) and std::partial_sort(first, last, last) (the heap-sort branch the
introsort falls into past its depth limit); stored: the keys and the resulting permutation.

    python tests/golden/make_std_sort_vectors.py        # needs g++; writes tests/golden/std_sort_vectors.npz
"""
import os
import subprocess
import tempfile

import numpy as np

SRC = r"""
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct E { float k; int id; };
int main(int argc, char** argv) {
  const int n = atoi(argv[1]), mod = atoi(argv[3]), heap = atoi(argv[4]); srand(atoi(argv[2]));
  std::vector<E> v(n);
  for (int i = 0; i < n; ++i) v[i] = E{(float)(rand() % mod), i};
  for (auto& e : v) printf("%g ", e.k); printf("\n");
  auto cmp = [](const E& a, const E& b) { return a.k < b.k; };
  if (heap) std::partial_sort(v.begin(), v.end(), v.end(), cmp); else std::sort(v.begin(), v.end(), cmp);
  for (auto& e : v) printf("%d ", e.id); printf("\n");
}
"""

if __name__ == "__main__":
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        with open(os.path.join(tmp, "s.cpp"), "w") as f:
            f.write(SRC)
        exe = os.path.join(tmp, "s")
        subprocess.run(["g++", "-O2", "-o", exe, os.path.join(tmp, "s.cpp")], check=True)
        for heap in (0, 1):
            for n in (2, 5, 16, 17, 33, 100, 257, 1000):
                for seed in (1, 2):
                    for mod in (2, 5, 1000000):
                        r = subprocess.run([exe, str(n), str(seed), str(mod), str(heap)], capture_output=True, text=True, check=True)
                        a, b = r.stdout.strip().split("\n")
                        key = f"{'heap' if heap else 'sort'}_{n}_{seed}_{mod}"
                        out[key + "/keys"] = np.array(a.split(), np.float32)
                        out[key + "/perm"] = np.array(b.split(), np.int32)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "std_sort_vectors.npz"), **out)
    print(len(out) // 2, "vectors")
