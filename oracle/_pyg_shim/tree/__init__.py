"""TEST INFRASTRUCTURE -- minimal stand-in for the `tree` (dm-tree) package, which the reference's vendored
openfold/np/residue_constants.py imports for one module-level `tree.map_structure` call over nested lists."""


def map_structure(fn, structure):
    if isinstance(structure, dict):
        return {k: map_structure(fn, v) for k, v in structure.items()}
    if isinstance(structure, (list, tuple)):
        return type(structure)(map_structure(fn, v) for v in structure)
    return fn(structure)
