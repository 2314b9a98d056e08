// TDirectDbWriter: fills the `assets` table of a FRESH afec-ll.db by writing sqlite's file format directly (table b-tree
// leaves with overflow chains, the filename index, interior pages, header), sequentially, with one copy of every row.
//
// Why: a row of the reference's schema (SqliteSampleDescriptorPool.cpp:1313-1350: 461 columns, 122 msgpack BLOBs) is
// 0.2-1 MB, i.e. 50-250 overflow pages; sqlite's insert path moves such rows at 0.5-0.7 GB/s (page cache, overflow chain
// bookkeeping, one pwrite per page) -- two orders of magnitude below what the GPU path delivers.
// The format itself is simple for an append-only load: rows arrive in rowid order, every row's cell goes to the current
// leaf and the rest of its payload to consecutive overflow pages.  The schema (page 1, the two root pages) is created by
// sqlite itself; this class only appends pages, rebuilds the two roots and updates the page count in the header, so
// every byte sqlite parses to find its way round was either written by sqlite or is covered by PRAGMA integrity_check
// (tests/test_host_sink.py).  The finished file is an ordinary database: the pool reopens it through sqlite.
#pragma once

#include <cstdint>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

namespace afec {

struct TDbValue {
  enum Kind { kNull, kInt, kReal, kText, kBlob } mKind = kNull;
  long long mInt = 0;
  double mReal = 0.0;
  const void* mData = nullptr;    // text / blob bytes: must stay valid until AddRow returns
  size_t mSize = 0;
  static TDbValue Null() { return TDbValue(); }
  static TDbValue Int(long long v) { TDbValue x; x.mKind = kInt; x.mInt = v; return x; }
  static TDbValue Real(double v) { TDbValue x; x.mKind = kReal; x.mReal = v; return x; }
  static TDbValue Text(const std::string& s) { TDbValue x; x.mKind = kText; x.mData = s.data(); x.mSize = s.size(); return x; }
  static TDbValue Text(const char* s, size_t n) { TDbValue x; x.mKind = kText; x.mData = s; x.mSize = n; return x; }
  static TDbValue Blob(const void* p, size_t n) { TDbValue x; x.mKind = kBlob; x.mData = p; x.mSize = n; return x; }
};

class TDirectDbWriter {
public:
  // `FileName`: a database that sqlite created and closed (no -wal / -journal beside it), holding the empty table whose
  // root page is `TableRoot` and its primary-key index (text key = column 0, then the rowid) with root `IndexRoot`
  TDirectDbWriter(const std::string& FileName, uint32_t TableRoot, uint32_t IndexRoot);
  ~TDirectDbWriter();
  // one row, all columns in table order; column 0 is the TEXT primary key (a key seen before throws: see HasKey -- the pool
  // keeps such rows aside and replaces through sqlite afterwards)
  void AddRow(const std::vector<TDbValue>& Values);
  long long Rows() const { return mRowId; }
  bool HasKey(const std::string& Key) const { return mKeys.count(Key) != 0; }
  // writes the interior pages, both roots and the header; the file is complete and closed afterwards
  void Finish();

private:
  struct Segment { const unsigned char* p; size_t n; };
  uint32_t LocalSize(uint64_t Payload, uint32_t MaxLocal) const;
  unsigned char* NextPageSlot(uint32_t& PageNo);
  uint32_t AppendPage(const unsigned char* Page);           // returns the page's number
  void WritePageAt(uint32_t PageNo, const unsigned char* Page);
  void FlushAppend();
  // copies the payload: the first `local` bytes to `Cell`, the rest to fresh overflow pages; returns the first overflow page (0: none)
  uint32_t SpillPayload(const std::vector<Segment>& Segs, uint64_t Payload, uint32_t Local, std::string& Cell);
  void FlushLeaf(bool ToRoot);
  static void BuildPage(std::vector<unsigned char>& Page, uint32_t PageSize, unsigned char Type, const std::vector<std::string>& Cells, uint32_t RightMost);
  void FinishTable();
  void FinishIndex();

  int mFd = -1;
  uint32_t mPageSize = 0, mUsable = 0, mNextPage = 0, mTableRoot = 0, mIndexRoot = 0;
  std::vector<unsigned char> mAppend;         // pages mAppendFirst .. (mFill bytes) waiting for one sequential write
  size_t mFill = 0;
  uint32_t mAppendFirst = 0;
  std::vector<std::string> mLeafCells; size_t mLeafBytes = 0; long long mLeafLastRow = 0; bool mLeafFlushed = false;
  std::vector<std::pair<uint32_t, long long>> mTableChildren;          // (leaf page, largest rowid in it)
  std::vector<std::pair<std::string, long long>> mIndex;               // (key text, rowid)
  std::unordered_set<std::string> mKeys;
  std::vector<unsigned char> mHeaderBuf, mScratch;
  long long mRowId = 0;
  bool mFinished = false;
};

}  // namespace afec
