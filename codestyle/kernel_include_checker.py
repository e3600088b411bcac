"""Kernel sources stay free of torch / ATen headers (they compile in seconds with plain nvcc and can be reused outside PyTorch);
only ``bindings.cpp`` may include them."""
import re
import sys

BAD = re.compile(r'#include\s*[<"](torch/|ATen/|c10/)')

if __name__ == "__main__":
    rc = 0
    for path in sys.argv[1:]:
        with open(path, encoding="utf-8") as f:
            for n, line in enumerate(f, 1):
                if BAD.search(line):
                    print(f"{path}:{n}: torch header in kernel source")
                    rc = 1
    sys.exit(rc)
