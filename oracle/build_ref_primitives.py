#!/usr/bin/env python
"""Builds oracle/_ref/libpumipic_ref_primitives.so: the REFERENCE'S OWN search code -- geometric
primitives and the search loops -- compiled unmodified from the sources where they lie under
/root/reference.  TEST INFRASTRUCTURE.

The reference as a whole is unbuildable here (Kokkos, Omega_h, EnGPar, MPI are absent; DESIGN.md
section 2), but its search is free functions and templates over a small vocabulary.  This script
  1. locates each function in the reference tree by its signature and copies its text (brace
     matching that skips comments and literals) into a temporary directory that is deleted
     after the compile: no reference source enters the repository or stays in the tree, only the
     shared library lands in the git-ignored oracle/_ref/,
  2. compiles it with g++, -ffp-contract=off like the oracle, against oracle/ref_shim/: our
     stand-ins for Omega_h's small-vector types and arithmetic (omega_h_shim.hpp), for
     Omega_h::Mesh / Write / Read, ps::parallel_for and the Kokkos / MPI calls the loops make
     (omega_h_mesh_shim.hpp: serial, the kernels being data-parallel over slots), and extern "C"
     wrappers with the oracle's signatures (ref_primitives.cpp).
tests/test_oracle_vs_reference_source.py then requires the oracle's restatements -- every
primitive and all four searches -- to agree with the reference's own code bit for bit.
The script is a no-op (exit 0) where /root/reference does not exist (the GPU box: the prebuilt
library travels with the snapshot).
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("PUMIPIC_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libpumipic_ref_primitives.so")

# (file, regex of the line that starts the definition, occurrence index)
FUNCTIONS = [
    ("src/pumipic_utils.hpp", r"bool all_positive\(const Vec a", 0),
    ("src/pumipic_utils.hpp", r"Omega_h::LO min3\(", 0),
    ("src/pumipic_utils.hpp", r"Omega_h::LO min_index\(const T a", 0),
    ("src/pumipic_utils.hpp", r"Omega_h::LO max_index\(const T a", 0),
    ("src/pumipic_utils.hpp", r"o::LO getFaceMap\(", 0),
    ("src/pumipic_utils.hpp", r"bool isFaceFlipped\(const o::LO ei", 0),
    ("src/pumipic_utils.hpp", r"bool isFaceFlipped\(const o::LO fi", 0),
    ("src/pumipic_utils.hpp", r"void get_face_from_face_index_of_tet\(", 0),
    ("src/pumipic_adjacency.hpp", r"bool find_barycentric_tet\(", 0),
    ("src/pumipic_adjacency.hpp", r"bool find_barycentric_tri_simple\(", 0),
    ("src/pumipic_adjacency.hpp", r"bool line_triangle_intx_simple \(", 0),
    ("src/pumipic_adjacency.tpp", r"void barycentric_tri\(", 0),
    ("src/pumipic_adjacency.tpp", r"bool barycentric_tet\(", 0),
    ("src/pumipic_adjacency.tpp", r"bool ray_intersects_triangle\(", 0),
    ("src/pumipic_adjacency.tpp", r"bool line_segment_intersects_triangle\(", 0),
    ("src/pumipic_adjacency.tpp", r"bool line_edge_2d\(", 0),
    ("src/pumipic_adjacency.tpp", r"o::LO find_exit_face_bcc_3d\(", 0),
    # the search loops (compiled against ref_shim/omega_h_mesh_shim.hpp)
    ("src/pumipic_adjacency.hpp", r"o::Vector<3> makeVector3\(int pid", 0),
    ("src/pumipic_adjacency.hpp", r"o::Vector<2> makeVector2\(int pid", 0),
    ("src/pumipic_adjacency.tpp", r"o::LO check_initial_parents\(", 0),
    ("src/pumipic_adjacency.tpp", r"void find_exit_face\(", 0),
    ("src/pumipic_adjacency.tpp", r"void check_model_intersection\(", 0),
    ("src/pumipic_adjacency.tpp", r"void set_new_element\(", 0),
    ("src/pumipic_adjacency.tpp", r"o::Real compute_tolerance_from_area\(", 0),
    ("src/pumipic_adjacency.tpp", r"bool trace_particle_through_mesh\(", 0),
    ("src/pumipic_adjacency.tpp", r"struct RemoveParticleOnGeometricModelExit", 0),
    ("src/pumipic_adjacency.tpp", r"bool search_mesh\(o::Mesh& mesh", 0),
    # search_mesh_2d (legacy 2D walk) with the 5-argument barycentric_tri it calls
    ("src/pumipic_adjacency.hpp", r"OMEGA_H_DEVICE void barycentric_tri\(", 0),
    ("src/pumipic_adjacency.hpp", r"bool search_mesh_2d\(o::Mesh& mesh", 0),
    # legacy 3D search_mesh and search_mesh_3d with their helpers
    ("src/pumipic_adjacency.hpp", r"o::Matrix<3, 3> gatherVectors3x3\(", 0),
    ("src/pumipic_adjacency.hpp", r"o::Matrix<3, 4> gatherVectors4x3\(", 0),
    ("src/pumipic_adjacency.hpp", r"bool barycentric_coords_tet\(", 0),
    ("src/pumipic_adjacency.hpp", r"OMEGA_H_DEVICE bool isPointWithinElemTet\(", 0),
    ("src/pumipic_adjacency.hpp", r"OMEGA_H_DEVICE bool isPointWithinElemTet\(", 1),
    ("src/pumipic_adjacency.hpp", r"bool search_mesh_3d\(o::Mesh& mesh", 0),
    ("src/pumipic_adjacency.hpp", r"bool search_mesh\(o::Mesh& mesh, ParticleStructure< ParticleType >\* ptcls", 0),
    # gather (field -> particle): tet vertex fields and the regular-grid interpolators
    ("src/pumipic_adjacency.hpp", r"Omega_h::Real interpolateTetVtx\(", 0),
    ("src/pumipic_adjacency.hpp", r"void interpolate3dFieldTet\(", 0),
    ("src/pumipic_adjacency.hpp", r"void findBCCoordsInTet\(", 0),
    ("src/pumipic_utils.hpp", r"o::Real interpolate2d_base\(", 0),
    ("src/pumipic_utils.hpp", r"o::Real interpolate2d_baseg\(", 0),
    ("src/pumipic_utils.hpp", r"o::Real interpolate2d_based\(", 0),
    ("src/pumipic_utils.hpp", r"o::Real interpolate2d\(const o::Reals& data, const o::Real gridXi", 0),
    ("src/pumipic_utils.hpp", r"o::Real interpolate2d_field\(", 0),
    ("src/pumipic_utils.hpp", r"o::Real interpolate3d_field\(", 0),
    ("src/pumipic_utils.hpp", r"void interp2dVector \(", 0),
]


# test/gyroScatter.hpp (global namespace; compiled in ref_shim/ref_xgcm.cpp against xgcm_shim.hpp)
XGCM_FUNCTIONS = [
    ("test/gyroScatter.hpp", r"namespace \{\n  o::Real gyro_rmax", 0),
    ("test/gyroScatter.hpp", r"void setGyroConfig\(", 0),
    ("test/gyroScatter.hpp", r"o::LOs searchAndBuildMap\(", 0),
    ("test/gyroScatter.hpp", r"void createGyroRingMappings\(", 0),
    ("test/gyroScatter.hpp", r"void gyroScatter\(", 0),
    ("test/ellipticalPush.hpp", r"namespace ellipticalPush \{", 0),
]
# test/test_adj.cpp: the generator of the synthetic inputs (compiled in ref_shim/ref_testadj.cpp)
TESTADJ_LINES = [r"#define PARTICLE_SEED [^\n]*", r"typedef p::MemberTypes<[^;]*> Particle;",
                 r"typedef p::ParticleStructure<Particle> PS;"]
TESTADJ_FUNCTIONS = [
    ("test/test_adj.cpp", r"int setSourceElements\(", 0),
    ("test/test_adj.cpp", r"void init2DInternal\(", 0),
    ("test/test_adj.cpp", r"void init3DInternal\(", 0),
    ("test/test_adj.cpp", r"o::Real determine_distance\(", 0),
    ("test/test_adj.cpp", r"o::Real get_push_distance\(", 0),
    ("test/test_adj.cpp", r"void push_ptcls\(", 0),
]
# test/pseudoPushAndSearch.cpp: the constant-vector push and the position update (ref_shim/ref_ppas.cpp)
PPAS_LINES = [r"typedef MemberTypes<Vector3d, Vector3d, int> Particle;", r"typedef ps::ParticleStructure<Particle> PS;"]
PPAS_FUNCTIONS = [
    ("test/pseudoPushAndSearch.cpp", r"void push\(PS\* ptcls, int np, fp_t distance", 0),
    ("test/pseudoPushAndSearch.cpp", r"void updatePtclPositions\(PS\* ptcls\)", 0),
]
# test/pseudoXGCm.cpp: the generator of the poloidal-plane particle load (ref_shim/ref_xgcm_init.cpp)
XGCINIT_LINES = [r"#define ELEMENT_SEED [^\n]*", r"#define PARTICLE_SEED [^\n]*"]
XGCINIT_TYPE_LINES = [r"typedef MemberTypes<Vector3d, Vector3d, int, float, float> Particle;",
                      r"typedef ps::ParticleStructure<Particle> PS;"]
XGCINIT_FUNCTIONS = [
    ("test/pseudoXGCm.cpp", r"int setSourceElements\(p::Mesh& picparts", 0),
    ("test/pseudoXGCm.cpp", r"void setInitialPtclCoords\(p::Mesh& picparts", 0),
]
# src/pumipic_part_construct.cpp: the set-up kernels of PICpart construction (ref_shim/ref_picpart.cpp);
# index 1 = the definition (index 0 is the forward declaration at the top of the file)
PICPART_FUNCTIONS = [
    ("src/pumipic_part_construct.cpp", r"void setOwnerByClassification\(Omega_h::Mesh& m", 1),
    ("src/pumipic_part_construct.cpp", r"Omega_h::LOs defineOwners\(Omega_h::Mesh& m", 1),
    ("src/pumipic_part_construct.cpp", r"Omega_h::LOs calculateOwnerOffset\(Omega_h::LOs owner", 1),
    ("src/pumipic_part_construct.cpp", r"struct GlobalNumberer \{", 0),
    ("src/pumipic_part_construct.cpp", r"Omega_h::LOs createGlobalNumbering\(Omega_h::LOs owner", 1),
    ("src/pumipic_part_construct.cpp", r"Omega_h::LOs rankLidNumbering\(Omega_h::LOs owner", 1),
    ("src/pumipic_part_construct.cpp", r"void BFS\(int nents", 0),
    ("src/pumipic_part_construct.cpp", r"void bfsBufferLayers\(Omega_h::Mesh& mesh", 1),
    ("src/pumipic_part_construct.cpp", r"void bfsSafeInward\(Omega_h::Mesh& mesh", 1),
]
# particle_structs/src/scs/SCS_buildFns.h: the Sell-C-sigma geometry (ref_shim/ref_scs.cpp)
SCS_FUNCTIONS = [
    ("particle_structs/src/scs/SCS_buildFns.h", r"int SellCSigma<DataTypes, MemSpace>::chooseChunkHeight\(", 0),
    ("particle_structs/src/scs/SCS_buildFns.h", r"void SellCSigma<DataTypes, MemSpace>::constructChunks\(", 0),
    ("particle_structs/src/scs/SCS_buildFns.h", r"void SellCSigma<DataTypes, MemSpace>::constructOffsets\(", 0),
]
# src/pumipic_ptcl_ops.hpp (namespace pumipic; needs the pumipic::Mesh stand-in of xgcm_shim.hpp)
PTCL_OPS_FUNCTIONS = [
    ("src/pumipic_ptcl_ops.hpp", r"void setUnsafeProcs\(Mesh& mesh", 0),
]
XGCM_TYPEDEFS = [r"typedef MemberTypes<[^;]*> Point;", r"typedef ps::ParticleStructure<Point> PSpt;",
                 r"typedef MemberTypes<[^;]*> Particle;", r"typedef ps::ParticleStructure<Particle> PS;"]


def extract(text, pattern, which):
    hits = [m for m in re.finditer(pattern, text)]
    if len(hits) <= which:
        raise SystemExit("build_ref_primitives: %r not found in the reference" % pattern)
    pos = hits[which].start()
    start = text.rfind("\n", 0, pos) + 1
    # include a preceding `template <...>` header (one line, or a few lines ending in '>')
    probe = start
    for _ in range(4):
        prev_start = text.rfind("\n", 0, probe - 1) + 1
        line = text[prev_start:probe].strip()
        if line.startswith("template"):
            start = prev_start
            break
        if not line.endswith(">") and not line.endswith(","):
            break
        probe = prev_start
    brace = text.index("{", pos)
    depth, i = 0, brace
    while True:                                       # brace matching that skips comments and literals
        c = text[i]
        if text.startswith("//", i):
            i = text.index("\n", i)
            continue
        if text.startswith("/*", i):
            i = text.index("*/", i) + 2
            continue
        if c == '"' or c == "'":
            j = i + 1
            while text[j] != c:
                j += 2 if text[j] == "\\" else 1
            i = j + 1
            continue
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        i += 1
    if text[i + 1:i + 40].lstrip().startswith(";"):      # struct / class definitions
        i = text.index(";", i)
    line0 = text.count("\n", 0, start) + 1
    line1 = text.count("\n", 0, i) + 1
    return text[start:i + 1], line0, line1


def main():
    if not os.path.isdir(REF):
        print("build_ref_primitives: %s not present, keeping the prebuilt library (if any)" % REF)
        return 0
    srcs = [os.path.join(REF, f) for f in sorted({f for f, _, _ in FUNCTIONS})]
    srcs += [os.path.join(REF, "src/pumipic_constants.hpp"), os.path.join(HERE, "ref_shim", "omega_h_shim.hpp"),
             os.path.join(HERE, "ref_shim", "omega_h_mesh_shim.hpp"),
             os.path.join(HERE, "ref_shim", "ref_primitives.cpp"), os.path.join(HERE, "ref_shim", "xgcm_shim.hpp"),
             os.path.join(HERE, "ref_shim", "ref_xgcm.cpp"), os.path.join(REF, "test/gyroScatter.hpp"),
             os.path.join(REF, "test/ellipticalPush.hpp"), os.path.join(REF, "src/pumipic_ptcl_ops.hpp"),
             os.path.join(REF, "test/test_adj.cpp"), os.path.join(HERE, "ref_shim", "ref_testadj.cpp"),
             os.path.join(REF, "test/pseudoPushAndSearch.cpp"), os.path.join(HERE, "ref_shim", "ref_ppas.cpp"),
             os.path.join(REF, "particle_structs/src/scs/SCS_buildFns.h"), os.path.join(HERE, "ref_shim", "ref_scs.cpp"),
             os.path.join(REF, "test/pseudoXGCm.cpp"), os.path.join(HERE, "ref_shim", "ref_xgcm_init.cpp"),
             os.path.join(REF, "src/pumipic_part_construct.cpp"), os.path.join(HERE, "ref_shim", "ref_picpart.cpp"),
             os.path.abspath(__file__)]
    if os.path.exists(LIB) and all(os.path.getmtime(s) <= os.path.getmtime(LIB) for s in srcs):
        return 0
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="pumipic_ref_")     # the extracted reference text never stays on disk
    parts = ["// GENERATED by oracle/build_ref_primitives.py from %s -- reference source text in a temporary\n"
             "// directory, deleted after the compile.\n" % REF]
    consts = open(os.path.join(REF, "src/pumipic_constants.hpp")).read()
    for name in ("EPSILON", "DIM", "FDIM"):
        m = re.search(r"^.*\b%s\b\s*=.*;.*$" % name, consts, re.M)
        if not m:
            raise SystemExit("build_ref_primitives: constant %s not found" % name)
        parts.append("// src/pumipic_constants.hpp\n" + m.group(0).strip() + "\n")
    cache = {}
    for f, pat, which in FUNCTIONS:
        text = cache.setdefault(f, open(os.path.join(REF, f)).read())
        body, l0, l1 = extract(text, pat, which)
        parts.append("// %s:%d-%d\n%s\n" % (f, l0, l1, body))
    with open(os.path.join(tmp, "ref_primitives.inc"), "w") as fh:
        fh.write("\n".join(parts))
    xparts = [parts[0]]
    types = open(os.path.join(REF, "test/pseudoXGCmTypes.hpp")).read()
    for pat in XGCM_TYPEDEFS:
        m = re.search(pat, types)
        if not m:
            raise SystemExit("build_ref_primitives: %r not found in pseudoXGCmTypes.hpp" % pat)
        xparts.append("// test/pseudoXGCmTypes.hpp\n" + m.group(0) + "\n")
    for f, pat, which in XGCM_FUNCTIONS:
        text = cache.setdefault(f, open(os.path.join(REF, f)).read())
        body, l0, l1 = extract(text, pat, which)
        xparts.append("// %s:%d-%d\n%s\n" % (f, l0, l1, body))
    with open(os.path.join(tmp, "ref_xgcm.inc"), "w") as fh:
        fh.write("\n".join(xparts))
    oparts = [parts[0]]
    for f, pat, which in PTCL_OPS_FUNCTIONS:
        text = cache.setdefault(f, open(os.path.join(REF, f)).read())
        body, l0, l1 = extract(text, pat, which)
        oparts.append("// %s:%d-%d\n%s\n" % (f, l0, l1, body))
    with open(os.path.join(tmp, "ref_ptcl_ops.inc"), "w") as fh:
        fh.write("\n".join(oparts))
    tparts = [parts[0]]
    tadj = open(os.path.join(REF, "test/test_adj.cpp")).read()
    for pat in TESTADJ_LINES:
        m = re.search(pat, tadj)
        if not m:
            raise SystemExit("build_ref_primitives: %r not found in test_adj.cpp" % pat)
        tparts.append("// test/test_adj.cpp\n" + m.group(0) + "\n")
    for f, pat, which in TESTADJ_FUNCTIONS:
        body, l0, l1 = extract(tadj, pat, which)
        tparts.append("// %s:%d-%d\n%s\n" % (f, l0, l1, body))
    with open(os.path.join(tmp, "ref_testadj.inc"), "w") as fh:
        fh.write("\n".join(tparts))
    sparts = [parts[0]]
    for f, pat, which in SCS_FUNCTIONS:
        text = cache.setdefault(f, open(os.path.join(REF, f)).read())
        body, l0, l1 = extract(text, pat, which)
        sparts.append("// %s:%d-%d\n%s\n" % (f, l0, l1, body))
    with open(os.path.join(tmp, "ref_scs.inc"), "w") as fh:
        fh.write("\n".join(sparts))
    pparts = [parts[0]]
    ppas = open(os.path.join(REF, "test/pseudoPushAndSearch.cpp")).read()
    for pat in PPAS_LINES:
        m = re.search(pat, ppas)
        if not m:
            raise SystemExit("build_ref_primitives: %r not found in pseudoPushAndSearch.cpp" % pat)
        pparts.append("// test/pseudoPushAndSearch.cpp\n" + m.group(0) + "\n")
    for f, pat, which in PPAS_FUNCTIONS:
        body, l0, l1 = extract(ppas, pat, which)
        pparts.append("// %s:%d-%d\n%s\n" % (f, l0, l1, body))
    with open(os.path.join(tmp, "ref_ppas.inc"), "w") as fh:
        fh.write("\n".join(pparts))
    cparts = [parts[0]]
    for f, pat, which in PICPART_FUNCTIONS:
        text = cache.setdefault(f, open(os.path.join(REF, f)).read())
        body, l0, l1 = extract(text, pat, which)
        cparts.append("// %s:%d-%d\n%s\n" % (f, l0, l1, body))
    with open(os.path.join(tmp, "ref_picpart.inc"), "w") as fh:
        fh.write("\n".join(cparts))
    iparts = [parts[0]]
    xgcm = open(os.path.join(REF, "test/pseudoXGCm.cpp")).read()
    xtypes = open(os.path.join(REF, "test/pseudoXGCmTypes.hpp")).read()
    for text, pats, name in ((xgcm, XGCINIT_LINES, "test/pseudoXGCm.cpp"),
                             (xtypes, XGCINIT_TYPE_LINES, "test/pseudoXGCmTypes.hpp")):
        for pat in pats:
            m = re.search(pat, text)
            if not m:
                raise SystemExit("build_ref_primitives: %r not found in %s" % (pat, name))
            iparts.append("// %s\n%s\n" % (name, m.group(0)))
    for f, pat, which in XGCINIT_FUNCTIONS:
        body, l0, l1 = extract(xgcm, pat, which)
        iparts.append("// %s:%d-%d\n%s\n" % (f, l0, l1, body))
    with open(os.path.join(tmp, "ref_xgcm_init.inc"), "w") as fh:
        fh.write("\n".join(iparts))
    cmd = ["g++", "-O3", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-std=c++17", "-Wno-unused-function",
           "-Wno-deprecated-declarations", "-I", os.path.join(HERE, "ref_shim"), "-I", tmp,
           os.path.join(HERE, "ref_shim", "ref_primitives.cpp"), os.path.join(HERE, "ref_shim", "ref_xgcm.cpp"),
           os.path.join(HERE, "ref_shim", "ref_testadj.cpp"), os.path.join(HERE, "ref_shim", "ref_ppas.cpp"),
           os.path.join(HERE, "ref_shim", "ref_scs.cpp"), os.path.join(HERE, "ref_shim", "ref_xgcm_init.cpp"),
           os.path.join(HERE, "ref_shim", "ref_picpart.cpp"), "-o", LIB]
    try:
        subprocess.check_call(cmd)
    finally:
        if os.environ.get("PUMIPIC_KEEP_REF_TEXT"):
            print("extracted reference text kept in", tmp)
        else:
            shutil.rmtree(tmp, ignore_errors=True)
    for stale in ("ref_primitives.inc", "ref_xgcm.inc", "ref_ptcl_ops.inc", "ref_testadj.inc", "ref_ppas.inc", "ref_scs.inc"):
        if os.path.exists(os.path.join(OUT, stale)):
            os.remove(os.path.join(OUT, stale))
    print(LIB)
    return 0


if __name__ == "__main__":
    sys.exit(main())
