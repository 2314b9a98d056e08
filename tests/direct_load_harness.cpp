// Test harness (tests/test_host_sink.py::test_direct_load_under_sanitizers): a direct load through the pool -- host-packed rows,
// failed rows, one name twice, names of a given length in unsorted order -- built with -fsanitize=address,undefined against
// sqlite_pool.cpp / direct_db_writer.cpp / descriptors.cpp alone (no CUDA library).   harness <db> <rows> <frames> <name length>
#include "afx_host.h"
#include <cstdio>
#include <cstring>
#include <vector>
#include <string>
using namespace afec;
int main(int argc, char** argv)
{
  const char* db = argv[1]; const int n = atoi(argv[2]); const int frames = atoi(argv[3]); const int name_len = atoi(argv[4]);
  TSampleDescriptors d;
  d.mFileType = "wav"; d.mFrames = frames; d.mRhythmFrames = frames * 8;
  for (int s = 0; s < AFX_N_FS; ++s) { const int m = s < AFX_N_FS_MAIN ? frames : frames * 8; d.mFramedScalars[s].resize(m); for (int i = 0; i < m; ++i) d.mFramedScalars[s][i] = 0.001 * i + s; }
  for (int v = 0; v < AFX_N_FV; ++v) { const size_t m = (size_t)frames * kFramedVectorBands[v]; d.mFramedVectors[v].resize(m); for (size_t i = 0; i < m; ++i) d.mFramedVectors[v][i] = 1e-3 * (double)i; }
  for (int s = 0; s < AFX_N_SERIES; ++s) for (int k = 0; k < AFX_N_STATS; ++k) d.mStats[s][k] = s + 0.01 * k;
  TSqliteSampleDescriptorPool pool;
  if (!pool.Open(db)) return 1;
  if (!pool.BeginDirectLoad()) return 2;
  for (int i = 0; i < n; ++i) {
    char tail[32]; snprintf(tail, sizeof(tail), "%07d.wav", (i * 7919) % n);
    std::string name = "/nonexistent/" + std::string((size_t)std::max(0, name_len - 24), (char)('a' + i % 3)) + tail;
    if (i % 4 == 3) pool.InsertFailedSample(name, "Sample failed to load: test");
    else { d.mFileName = name; pool.InsertSample(name, d); }
    if (i == n / 2) pool.InsertFailedSample(name, "twice");       // a duplicate: deferred
  }
  pool.EndBulkLoad();
  printf("rows %d\n", pool.NumberOfSamples());
  pool.Close();
  return 0;
}
namespace afec { int ModificationStatTime(const std::string&) { return 0; } std::string ExtractFileExtension(const std::string& f) { return "wav"; } }
