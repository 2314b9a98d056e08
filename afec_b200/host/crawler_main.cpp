// afec-b200-crawler: a minimal stand-in for `Crawler -l low -o afec-ll.db <paths...>` (Crawler.cpp:136-386,
// 566-760) that drives the GPU path: collect files, build the change list against the database's modtimes
// (Crawler.cpp:934-998), analyse in batches on the listed devices, write afec-ll.db.
// Only the low-level set is produced (classification stays with the reference's host tools) and only WAV and
// AIFF are decoded here (FLAC / Ogg / MP3 decoding is the reference's CoreFileFormats, outside this path).
#include "afx_host.h"

#include <algorithm>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <chrono>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <unistd.h>
#include <sys/stat.h>

using namespace afec;

static volatile bool sAbort = false;
static void on_sigint(int) { sAbort = true; }

static bool is_dir(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
static bool has_audio_ext(const std::string& p) { return IsSupportedAudioFileExtension(p); }
static void collect(const std::string& path, std::vector<std::string>& out)
{
  if (!is_dir(path)) { if (has_audio_ext(path)) out.push_back(path); return; }
  DIR* d = opendir(path.c_str());
  if (!d) return;
  std::vector<std::string> names;
  while (dirent* e = readdir(d)) { if (e->d_name[0] != '.') names.push_back(e->d_name); }
  closedir(d);
  std::sort(names.begin(), names.end());
  for (const auto& n : names) collect(path + (path.back() == '/' ? "" : "/") + n, out);
}
static std::string abs_path(const std::string& p)
{
  char buf[4096];
  return realpath(p.c_str(), buf) ? std::string(buf) : p;
}

int main(int argc, char** argv)
{
  std::string out_db = "afec-ll.db", level = "low";
  int hop = 1024, slots = 3, shards = 1, decode_threads = 0; std::vector<int> devices(1, 0); std::vector<std::string> paths;
  bool quiet = false, merge = false, host_pack = false, direct_load = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return argv[++i]; };
    if (a == "-o" || a == "--out") out_db = next();
    else if (a == "-l" || a == "--level") level = next();
    else if (a == "-j" || a == "--jobs") slots = std::max(1, atoi(next().c_str()));
    else if (a == "--hop") hop = atoi(next().c_str());
    else if (a == "--devices") { devices.clear(); std::string s = next(); size_t p = 0; while (p <= s.size()) { size_t q = s.find(',', p); if (q == std::string::npos) q = s.size(); if (q > p) devices.push_back(atoi(s.substr(p, q - p).c_str())); p = q + 1; } }
    else if (a == "--shards") shards = std::max(1, atoi(next().c_str()));
    else if (a == "--merge") merge = true;
    else if (a == "--decode-threads") decode_threads = std::max(1, atoi(next().c_str()));
    else if (a == "--host-pack") host_pack = true;
    else if (a == "--page-size") setenv("AFX_SINK_PAGE_SIZE", next().c_str(), 1);
    else if (a == "--direct-load") direct_load = true;
    else if (a == "-q") quiet = true;
    else if (a == "-h" || a == "--help") {
      printf("usage: %s [-l low] [-o afec-ll.db] [-j slots-per-gpu] [--hop 1024] [--devices 0,1,..] [--decode-threads N]\n"
             "          [--shards N [--merge]] [--host-pack] [--page-size N] [--direct-load] <file-or-dir>...\n"
             "  --shards N   N sqlite writers side by side: afec-ll.db, afec-ll.db.1 .. .N-1 (each a valid afec-ll.db holding a\n"
             "               disjoint part of the rows); --merge appends the shards to afec-ll.db afterwards and deletes them\n"
             "  --host-pack  pack the msgpack BLOBs on the host instead of the GPU\n"
             "  --direct-load  a NEW database is written in sqlite's file format directly (no sqlite in the data path of the load)\n"
             "  --page-size N  sqlite page size of a NEW database (power of two, 512 .. 65536; default: sqlite's 4096 as the reference)\n", argv[0]);
      return 0;
    } else if (!a.empty() && a[0] == '-') { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    else paths.push_back(a);
  }
  if (level != "low") { fprintf(stderr, "only --level low runs on the GPU path (high-level classification stays with the reference Crawler)\n"); return 1; }
  if (paths.empty()) { fprintf(stderr, "no input paths\n"); return 1; }
  signal(SIGINT, on_sigint);
  try {
    std::vector<std::string> files;
    for (const auto& p : paths) collect(abs_path(p), files);
    TSqliteSampleDescriptorPool pool;
    const std::string db_abs = abs_path(out_db).empty() ? out_db : out_db;
    if (!pool.Open(db_abs)) { fprintf(stderr, "failed to open database %s\n", out_db.c_str()); return 1; }
    // file names are stored relative to the database's directory when every input lives below it
    // (SqliteSampleDescriptorPool.cpp:1164-1188; Crawler.cpp:610-640)
    {
      std::string dir = abs_path(out_db);
      const size_t s = dir.find_last_of('/');
      dir = (s == std::string::npos) ? abs_path(".") : dir.substr(0, s);
      if (dir.empty() || dir.back() != '/') dir += '/';
      bool all_below = !files.empty();
      for (const auto& f : files) if (f.compare(0, dir.size(), dir) != 0) { all_below = false; break; }
      if (all_below) pool.SetBasePath(dir);
    }
    // change list as SBuildChangeList (Crawler.cpp:933-998): a row is removed only when its file no longer exists on
    // disk, refreshed when the file's mtime is newer than the stored one -- whether or not this run's inputs name it --
    // and every input the database does not know is added
    std::vector<std::string> todo, gone;
    if (pool.IsEmpty()) todo = files;
    else {
      std::set<std::string> known;
      for (const auto& e : pool.SampleModificationDates()) {
        known.insert(e.first);
        struct stat st;
        if (stat(e.first.c_str(), &st) != 0) gone.push_back(e.first);
        else if (ModificationStatTime(e.first) > e.second) todo.push_back(e.first);
      }
      for (const auto& f : files) if (known.find(f) == known.end()) todo.push_back(f);
    }
    if (!gone.empty()) pool.RemoveSamples(gone);
    if (!quiet) printf("%zu files found, %zu to analyse, %zu removed\n", files.size(), todo.size(), gone.size());
    if (todo.empty()) return 0;
    TGpuSampleAnalyser analyser(44100, 2048, hop, devices, slots, !host_pack);
    if (decode_threads) analyser.SetDecodeThreads(decode_threads);
    // a fresh database is filled without a journal (TSqliteSampleDescriptorPool::BeginBulkLoad), shards always are fresh
    const bool bulk_main = (direct_load && pool.BeginDirectLoad()) || pool.BeginBulkLoad();
    std::vector<std::unique_ptr<TSqliteSampleDescriptorPool>> shard_pools;
    std::vector<std::string> shard_files;
    std::vector<TSampleDescriptorPool*> pools(1, &pool);
    std::vector<std::unique_ptr<std::mutex>> lock_store; std::vector<std::mutex*> locks;
    for (int k = 1; k < shards; ++k) {
      const std::string name = db_abs + "." + std::to_string(k);
      unlink(name.c_str()); unlink((name + "-wal").c_str()); unlink((name + "-shm").c_str());
      std::unique_ptr<TSqliteSampleDescriptorPool> sp(new TSqliteSampleDescriptorPool());
      if (!sp->Open(name)) { fprintf(stderr, "failed to open shard %s\n", name.c_str()); return 1; }
      sp->SetBasePath(pool.BasePath());
      if (!(direct_load && sp->BeginDirectLoad())) sp->BeginBulkLoad();
      pools.push_back(sp.get()); shard_files.push_back(name); shard_pools.push_back(std::move(sp));
    }
    for (size_t k = 0; k < pools.size(); ++k) { lock_store.emplace_back(new std::mutex()); locks.push_back(lock_store.back().get()); }
    TGpuSampleAnalyser::TProgress pr;
    const int failed = analyser.ExtractBatchSharded(todo, pools, locks, &pr, &sAbort);
    if (bulk_main) pool.EndBulkLoad();
    for (auto& sp : shard_pools) sp->Close();
    double merge_s = 0.0;
    if (merge && !shard_files.empty()) {
      const auto m0 = std::chrono::steady_clock::now();
      pool.MergeFrom(shard_files, true);
      merge_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - m0).count();
      if (!quiet) printf("merged %zu shards in %.3f s\n", shard_files.size(), merge_s);
    }
    if (!quiet) printf("{\"files\": %lld, \"failed\": %d, \"main_frames\": %lld, \"rhythm_frames\": %lld, \"audio_seconds\": %.3f, \"seconds\": %.4f, \"merge_seconds\": %.4f, \"shards\": %d, \"audio_hours_per_s\": %.4f}\n",
                       (long long)pr.mFiles, failed, (long long)pr.mMainFrames, (long long)pr.mRhythmFrames, pr.mAudioSeconds, pr.mSeconds, merge_s, shards,
                       pr.mSeconds + merge_s > 0 ? pr.mAudioSeconds / 3600.0 / (pr.mSeconds + merge_s) : 0.0);
    return sAbort ? 2 : 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
