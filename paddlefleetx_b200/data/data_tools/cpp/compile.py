"""Build the C++ sample-index helper (reference ppfleetx/data/data_tools/cpp/compile.py runs ``make``; here the extension is compiled
in-tree by ``ops/build.py`` with the same compiler flags as the rest of the native code)."""


def compile_helper(force: bool = False):
    """Compile ``fast_index_map_helpers``; call on ONE process per node (rank 0), then barrier."""
    from ....ops.build import build_data_helper

    return build_data_helper(force=force)


if __name__ == "__main__":
    print(compile_helper())
