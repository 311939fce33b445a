#include "mesh_data.h"

#include <sys/stat.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <iomanip>

#include "mesh_topology.h"

namespace mfb {

namespace {

inline uint64_t mix64 (uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// uniform in [-0.1, 0.1), a pure function of (global node, axis, seed)
inline double jitter (int64_t node, int axis, uint64_t seed)
{
    uint64_t h = mix64 (mix64 (seed) ^ (uint64_t)(node * 3 + axis));
    return ((double)(h >> 11) * (1.0 / 9007199254740992.0) - 0.5) * 0.2;
}

// Kuhn split: the six monotone lattice paths (0,0,0) -> (1,1,1)
const int kPaths[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};

struct Range { int lo, hi; };   // cubes [lo, hi)

inline Range block_range (int n, int parts, int b)
{
    return { (int)((int64_t)n * b / parts), (int)((int64_t)n * (b + 1) / parts) };
}

void make_dirs (const std::string &file)
{
    for (size_t p = file.find ('/', 1); p != std::string::npos; p = file.find ('/', p + 1)) {
        mkdir (file.substr (0, p).c_str (), 0777);
    }
}

}  // namespace

int generate_block (int nx, int ny, int nz, int px, int py, int pz, int rank,
                    uint64_t seed, SubMesh &out)
{
    if (nx < 1 || ny < 1 || nz < 1 || px < 1 || py < 1 || pz < 1) return -1;
    if (px > nx || py > ny || pz > nz || rank < 0 || rank >= px * py * pz) return -1;
    const int bx = rank % px, by = (rank / px) % py, bz = rank / (px * py);
    const Range rx = block_range (nx, px, bx), ry = block_range (ny, py, by),
                rz = block_range (nz, pz, bz);
    const int lx = rx.hi - rx.lo, ly = ry.hi - ry.lo, lz = rz.hi - rz.lo;
    const int64_t nodes64 = (int64_t)(lx + 1) * (ly + 1) * (lz + 1);
    const int64_t elems64 = (int64_t)lx * ly * lz * 6;
    if (nodes64 > INT32_MAX || elems64 * 16 > INT32_MAX) return -1;

    out = SubMesh ();
    out.nbNodes = (int)nodes64;
    out.nbElem = (int)elems64;
    out.coord.resize ((size_t)out.nbNodes * 3);
    out.boundNodesCode.resize (out.nbNodes);
    out.globalNode.resize (out.nbNodes);
    out.elemToNode.resize ((size_t)out.nbElem * 4);

    auto local_id = [&] (int i, int j, int k) {   // global lattice indices -> 1-based local id
        return ((k - rz.lo) * (ly + 1) + (j - ry.lo)) * (lx + 1) + (i - rx.lo) + 1;
    };

    for (int k = rz.lo; k <= rz.hi; k++) {
        for (int j = ry.lo; j <= ry.hi; j++) {
            for (int i = rx.lo; i <= rx.hi; i++) {
                const int n = local_id (i, j, k) - 1;
                const int64_t g = ((int64_t)k * (ny + 1) + j) * (nx + 1) + i;
                out.globalNode[n] = g;
                out.coord[(size_t)n * 3 + 0] = i + jitter (g, 0, seed);
                out.coord[(size_t)n * 3 + 1] = j + jitter (g, 1, seed);
                out.coord[(size_t)n * 3 + 2] = k + jitter (g, 2, seed);
                int code = 0;
                if (i == 0) code = 52;
                if (j == 0) code = 53;
                if (k == 0) code = 54;
                if (j == ny) code = 10;
                out.boundNodesCode[n] = code;
                if (code) out.nbBoundNodes++;
            }
        }
    }

    size_t e = 0;
    for (int k = rz.lo; k < rz.hi; k++) {
        for (int j = ry.lo; j < ry.hi; j++) {
            for (int i = rx.lo; i < rx.hi; i++) {
                for (int p = 0; p < 6; p++) {
                    int d[3] = {0, 0, 0};
                    out.elemToNode[e++] = local_id (i, j, k);
                    d[kPaths[p][0]] = 1;
                    out.elemToNode[e++] = local_id (i + d[0], j + d[1], k + d[2]);
                    d[kPaths[p][1]] = 1;
                    out.elemToNode[e++] = local_id (i + d[0], j + d[1], k + d[2]);
                    out.elemToNode[e++] = local_id (i + 1, j + 1, k + 1);
                }
            }
        }
    }

    // One interface per neighbouring block (face, edge or corner contact), in increasing
    // rank; nodes in increasing global id on both sides, so position j matches.
    struct Intf { int rank; std::vector<int> nodes; };
    std::vector<Intf> intfs;
    for (int dz = -1; dz <= 1; dz++) {
        for (int dy = -1; dy <= 1; dy++) {
            for (int dx = -1; dx <= 1; dx++) {
                if (!dx && !dy && !dz) continue;
                const int ox = bx + dx, oy = by + dy, oz = bz + dz;
                if (ox < 0 || ox >= px || oy < 0 || oy >= py || oz < 0 || oz >= pz) continue;
                const Range si = dx < 0 ? Range{rx.lo, rx.lo} : dx > 0 ? Range{rx.hi, rx.hi} : Range{rx.lo, rx.hi};
                const Range sj = dy < 0 ? Range{ry.lo, ry.lo} : dy > 0 ? Range{ry.hi, ry.hi} : Range{ry.lo, ry.hi};
                const Range sk = dz < 0 ? Range{rz.lo, rz.lo} : dz > 0 ? Range{rz.hi, rz.hi} : Range{rz.lo, rz.hi};
                Intf it;
                it.rank = (oz * py + oy) * px + ox;
                for (int k = sk.lo; k <= sk.hi; k++)
                    for (int j = sj.lo; j <= sj.hi; j++)
                        for (int i = si.lo; i <= si.hi; i++) it.nodes.push_back (local_id (i, j, k));
                intfs.push_back (std::move (it));
            }
        }
    }
    std::sort (intfs.begin (), intfs.end (), [] (const Intf &a, const Intf &b) { return a.rank < b.rank; });
    out.nbIntf = (int)intfs.size ();
    out.neighborsList.assign ((size_t)std::max (out.nbIntf, 1) * 3, 0);
    out.intfIndex.assign ((size_t)out.nbIntf + 1, 0);
    for (int q = 0; q < out.nbIntf; q++) {
        out.neighborsList[q] = intfs[q].rank + 1;
        out.intfIndex[q + 1] = out.intfIndex[q] + (int)intfs[q].nodes.size ();
        out.intfNodes.insert (out.intfNodes.end (), intfs[q].nodes.begin (), intfs[q].nodes.end ());
    }
    out.nbIntfNodes = (int)out.intfNodes.size ();

    const int64_t edges = count_csr_entries (out.elemToNode.data (), out.nbElem, out.nbNodes);
    if (edges * 9 > INT32_MAX) return -1;     // the reference indexes values with int
    out.nbEdges = (int)edges;
    return 0;
}

void choose_blocks (int nx, int ny, int nz, int maxRanks, int &px, int &py, int &pz)
{
    px = py = pz = 1;
    double bestScore = -1;
    for (int a = 1; a <= std::min (nx, maxRanks); a++) {
        for (int b = 1; b <= std::min (ny, maxRanks / a); b++) {
            for (int c = 1; c <= std::min (nz, maxRanks / (a * b)); c++) {
                const double sx = (double)nx / a, sy = (double)ny / b, sz = (double)nz / c;
                const double surface = sx * sy + sy * sz + sx * sz;
                // more ranks first, then the smallest block surface
                const double score = (double)a * b * c * 1e9 - surface;
                if (score > bestScore) { bestScore = score; px = a; py = b; pz = c; }
            }
        }
    }
}

std::string input_path (const std::string &dataPath, const std::string &mesh,
                        const std::string &op, int nbBlocks, int rank)
{
    return dataPath + "/" + mesh + "/inputs/" + op + "_" + std::to_string (nbBlocks) + "_" +
           std::to_string (rank);
}

std::string checking_path (const std::string &dataPath, const std::string &mesh,
                           const std::string &op, int nbBlocks, int rank)
{
    return dataPath + "/" + mesh + "/checkings/" + op + "_" + std::to_string (nbBlocks) + "_" +
           std::to_string (rank);
}

int write_input (const std::string &file, const SubMesh &m)
{
    make_dirs (file);
    FILE *f = fopen (file.c_str (), "wb");
    if (!f) return -1;
    const int header[6] = {m.nbElem, m.nbNodes, m.nbEdges, m.nbIntf, m.nbIntfNodes, m.nbBoundNodes};
    bool ok = fwrite (header, sizeof (int), 6, f) == 6;
    auto put = [&] (const void *p, size_t bytes) { ok = ok && (bytes == 0 || fwrite (p, 1, bytes, f) == bytes); };
    put (m.coord.data (), sizeof (double) * (size_t)m.nbNodes * 3);
    put (m.elemToNode.data (), sizeof (int) * (size_t)m.nbElem * 4);
    put (m.neighborsList.data (), sizeof (int) * (size_t)std::max (m.nbIntf, 1) * 3);
    put (m.intfIndex.data (), sizeof (int) * ((size_t)m.nbIntf + 1));
    put (m.intfNodes.data (), sizeof (int) * (size_t)m.nbIntfNodes);
    put (m.boundNodesCode.data (), sizeof (int) * (size_t)m.nbNodes);
    return (fclose (f) == 0 && ok) ? 0 : -1;
}

int read_input (const std::string &file, SubMesh &m)
{
    FILE *f = fopen (file.c_str (), "rb");
    if (!f) return -1;
    int header[6];
    bool ok = fread (header, sizeof (int), 6, f) == 6;
    if (ok) {
        for (int h : header) ok = ok && h >= 0;
    }
    if (!ok) { fclose (f); return -1; }
    m = SubMesh ();
    m.nbElem = header[0]; m.nbNodes = header[1]; m.nbEdges = header[2];
    m.nbIntf = header[3]; m.nbIntfNodes = header[4]; m.nbBoundNodes = header[5];
    m.coord.resize ((size_t)m.nbNodes * 3);
    m.elemToNode.resize ((size_t)m.nbElem * 4);
    m.neighborsList.resize ((size_t)std::max (m.nbIntf, 1) * 3);
    m.intfIndex.resize ((size_t)m.nbIntf + 1);
    m.intfNodes.resize ((size_t)m.nbIntfNodes);
    m.boundNodesCode.resize ((size_t)m.nbNodes);
    auto get = [&] (void *p, size_t bytes) { ok = ok && (bytes == 0 || fread (p, 1, bytes, f) == bytes); };
    get (m.coord.data (), sizeof (double) * m.coord.size ());
    get (m.elemToNode.data (), sizeof (int) * m.elemToNode.size ());
    get (m.neighborsList.data (), sizeof (int) * m.neighborsList.size ());
    get (m.intfIndex.data (), sizeof (int) * m.intfIndex.size ());
    get (m.intfNodes.data (), sizeof (int) * m.intfNodes.size ());
    get (m.boundNodesCode.data (), sizeof (int) * m.boundNodesCode.size ());
    fclose (f);
    return ok ? 0 : -1;
}

int write_checking (const std::string &file, double matrixNorm, double precNorm)
{
    make_dirs (file);
    std::ofstream out (file, std::ios::out | std::ios::trunc);
    if (!out.is_open ()) return -1;
    out << std::setprecision (17) << matrixNorm << std::endl << precNorm << std::endl;
    return out.good () ? 0 : -1;
}

int read_checking (const std::string &file, double &matrixNorm, double &precNorm)
{
    std::ifstream in (file, std::ios::in);
    if (!in.is_open ()) return -1;
    in >> matrixNorm >> precNorm;
    return in.fail () ? -1 : 0;
}

}  // namespace mfb
