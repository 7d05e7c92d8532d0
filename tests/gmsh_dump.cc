// tests/gmsh_dump.cc — test helper: reads a Gmsh file with the facade's GmshReader, refines it, and dumps the mesh as text
// (vertices, elements, physical groups, boundary segments) so that tests/test_gmsh.py can compare it entry by entry with an
// independent Python reading of the same file.  Uses no device call.
#include <cstdio>
#include <cstdlib>

#include "mdb200/multidomain.hh"

int main(int argc, char** argv)
{
  try {
    if (argc < 3) return 2;
    mdb200::GmshMesh m = mdb200::GmshReader::read(argv[1], false, true);
    m.globalRefine(std::atoi(argv[2]));
    std::printf("%lld %lld %lld\n", (long long)m.numVertices(), (long long)m.numElements(), (long long)m.boundaryIndexToPhysicalGroup.size());
    for (int64_t v = 0; v < m.numVertices(); ++v) std::printf("%.17g %.17g\n", m.vertices[(size_t)(2 * v)], m.vertices[(size_t)(2 * v + 1)]);
    for (int64_t e = 0; e < m.numElements(); ++e)
      std::printf("%lld %lld %lld %d\n", (long long)m.elements[(size_t)(3 * e)], (long long)m.elements[(size_t)(3 * e + 1)],
                  (long long)m.elements[(size_t)(3 * e + 2)], m.elementIndexToPhysicalGroup[(size_t)e]);
    for (size_t b = 0; b < m.boundaryIndexToPhysicalGroup.size(); ++b)
      std::printf("%lld %lld %d\n", (long long)m.boundarySegments[2 * b], (long long)m.boundarySegments[2 * b + 1], m.boundaryIndexToPhysicalGroup[b]);
    return 0;
  } catch (mdb200::Exception& e) {
    std::fprintf(stderr, "mdb200 error %d: %s\n", e.code, e.what());
    return 1;
  }
}
