// See direct_db_writer.h.  File-format facts used here (sqlite.org/fileformat2.html): the 100-byte header; b-tree page
// headers (8 bytes leaf / 12 bytes interior) followed by the 2-byte cell pointer array, cells packed at the end of the page;
// table leaf cell = varint(payload size) varint(rowid) payload [first overflow page]; table interior cell = left child,
// varint(key); index cells carry a record (key text, rowid) as payload; a payload larger than X (usable - 35 for table
// leaves, ((usable - 12) * 64 / 255) - 23 for index pages) keeps M + (P - M) % (usable - 4) bytes local (or M when that
// exceeds X; M = ((usable - 12) * 32 / 255) - 23) and chains the rest through overflow pages (next page number, then
// usable - 4 bytes); records = varint(header size), one serial type per column, the values; the page that holds byte
// 2^30 of the file (the lock-byte page) is never used.
#include "direct_db_writer.h"
#include "afx_host.h"

#include <algorithm>
#include <cstring>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

namespace afec {

namespace {

inline int varint_len(uint64_t v) { int n = 1; while (v >>= 7) ++n; return n; }          // v < 2^56: at most 8 bytes
inline void put_varint(std::string& o, uint64_t v)
{
  unsigned char b[10];
  const int n = varint_len(v);
  for (int i = n - 1; i >= 0; --i) { b[i] = (unsigned char)((v & 0x7f) | (i == n - 1 ? 0 : 0x80)); v >>= 7; }
  o.append((const char*)b, (size_t)n);
}
inline void put_be32(unsigned char* p, uint32_t v) { p[0] = (unsigned char)(v >> 24); p[1] = (unsigned char)(v >> 16); p[2] = (unsigned char)(v >> 8); p[3] = (unsigned char)v; }
inline void put_be16(unsigned char* p, uint32_t v) { p[0] = (unsigned char)(v >> 8); p[1] = (unsigned char)v; }
inline uint32_t get_be32(const unsigned char* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
inline uint32_t get_be16(const unsigned char* p) { return ((uint32_t)p[0] << 8) | p[1]; }
inline std::string be32_string(uint32_t v) { unsigned char b[4]; put_be32(b, v); return std::string((const char*)b, 4); }

// serial type of an integer and its big-endian two's complement bytes
inline int int_serial(long long v, unsigned char* out, int& bytes)
{
  if (v == 0) { bytes = 0; return 8; }
  if (v == 1) { bytes = 0; return 9; }
  int type;
  if (v >= -128 && v <= 127) { type = 1; bytes = 1; }
  else if (v >= -32768 && v <= 32767) { type = 2; bytes = 2; }
  else if (v >= -8388608 && v <= 8388607) { type = 3; bytes = 3; }
  else if (v >= -2147483648LL && v <= 2147483647LL) { type = 4; bytes = 4; }
  else if (v >= -140737488355328LL && v <= 140737488355327LL) { type = 5; bytes = 6; }
  else { type = 6; bytes = 8; }
  const uint64_t u = (uint64_t)v;
  for (int i = 0; i < bytes; ++i) out[i] = (unsigned char)(u >> (8 * (bytes - 1 - i)));
  return type;
}

void full_pwrite(int fd, const unsigned char* p, size_t n, off_t off)
{
  while (n) {
    const ssize_t w = pwrite(fd, p, n, off);
    if (w <= 0) throw TReadableException("direct database load: write failed");
    p += w; n -= (size_t)w; off += w;
  }
}

}  // namespace

TDirectDbWriter::TDirectDbWriter(const std::string& FileName, uint32_t TableRoot, uint32_t IndexRoot)
  : mTableRoot(TableRoot), mIndexRoot(IndexRoot)
{
  struct stat st;
  if (stat((FileName + "-wal").c_str(), &st) == 0 && st.st_size > 0) throw TReadableException("direct database load: a write-ahead log is present");
  mFd = open(FileName.c_str(), O_RDWR);
  if (mFd < 0) throw TReadableException("direct database load: cannot open " + FileName);
  mHeaderBuf.resize(100);
  if (pread(mFd, mHeaderBuf.data(), 100, 0) != 100 || memcmp(mHeaderBuf.data(), "SQLite format 3", 16) != 0) { close(mFd); mFd = -1; throw TReadableException("direct database load: not a database"); }
  mPageSize = get_be16(mHeaderBuf.data() + 16); if (mPageSize == 1) mPageSize = 65536;
  mUsable = mPageSize - mHeaderBuf[20];
  const uint32_t pages = get_be32(mHeaderBuf.data() + 28);
  // a file longer than its header says is what a load that was cut off leaves behind (the header is written last): the tail
  // is not part of the database and is written over
  const bool sane = mPageSize >= 512 && (mPageSize & (mPageSize - 1)) == 0 && mUsable >= 480 && fstat(mFd, &st) == 0 && pages >= 3 &&
                    (uint64_t)st.st_size >= (uint64_t)pages * mPageSize && get_be32(mHeaderBuf.data() + 52) == 0 &&          // no auto-vacuum: no pointer-map pages
                    get_be32(mHeaderBuf.data() + 56) == 1 && TableRoot >= 2 && IndexRoot >= 2 && TableRoot <= pages && IndexRoot <= pages;
  if (!sane) { close(mFd); mFd = -1; throw TReadableException("direct database load: unexpected database header"); }
  mNextPage = pages + 1;
  mAppendFirst = mNextPage;
  mAppend.resize(((size_t)16 << 20) / mPageSize * mPageSize);      // the append buffer: mFill bytes of it are waiting
}

TDirectDbWriter::~TDirectDbWriter() { if (mFd >= 0) close(mFd); }

uint32_t TDirectDbWriter::LocalSize(uint64_t Payload, uint32_t MaxLocal) const
{
  if (Payload <= MaxLocal) return (uint32_t)Payload;
  const uint32_t M = ((mUsable - 12) * 32 / 255) - 23;
  const uint32_t K = M + (uint32_t)((Payload - M) % (mUsable - 4));
  return K <= MaxLocal ? K : M;
}

void TDirectDbWriter::FlushAppend()
{
  if (!mFill) return;
  full_pwrite(mFd, mAppend.data(), mFill, (off_t)(mAppendFirst - 1) * (off_t)mPageSize);
  mAppendFirst += (uint32_t)(mFill / mPageSize);
  mFill = 0;
}

// the buffer slot of the next page of the file (mNextPage, which it advances); the lock-byte page is left blank and skipped.
// The slot is NOT cleared: the caller writes all of it.
unsigned char* TDirectDbWriter::NextPageSlot(uint32_t& PageNo)
{
  const uint32_t lock_page = (uint32_t)(0x40000000u / mPageSize) + 1;
  if (mNextPage == lock_page) {
    if (mFill == mAppend.size()) FlushAppend();
    memset(mAppend.data() + mFill, 0, mPageSize); mFill += mPageSize; ++mNextPage;
  }
  if (mFill == mAppend.size()) FlushAppend();
  unsigned char* slot = mAppend.data() + mFill;
  mFill += mPageSize;
  PageNo = mNextPage++;
  return slot;
}

uint32_t TDirectDbWriter::AppendPage(const unsigned char* Page)
{
  uint32_t no;
  memcpy(NextPageSlot(no), Page, mPageSize);
  return no;
}

void TDirectDbWriter::WritePageAt(uint32_t PageNo, const unsigned char* Page)
{
  full_pwrite(mFd, Page, mPageSize, (off_t)(PageNo - 1) * (off_t)mPageSize);
}

uint32_t TDirectDbWriter::SpillPayload(const std::vector<Segment>& Segs, uint64_t Payload, uint32_t Local, std::string& Cell)
{
  size_t si = 0, so = 0;                                         // cursor into the segments
  auto take = [&](unsigned char* dst, size_t n) {
    while (n) {
      const size_t m = std::min(n, Segs[si].n - so);
      memcpy(dst, Segs[si].p + so, m);
      dst += m; n -= m; so += m;
      if (so == Segs[si].n) { ++si; so = 0; }
    }
  };
  const size_t c0 = Cell.size();
  Cell.resize(c0 + Local);
  take((unsigned char*)&Cell[c0], Local);
  uint64_t rest = Payload - Local;
  if (rest == 0) return 0;
  const uint32_t lock_page = (uint32_t)(0x40000000u / mPageSize) + 1;
  const uint32_t per = mUsable - 4;
  uint32_t first = 0;
  while (rest) {
    uint32_t me;
    unsigned char* slot = NextPageSlot(me);
    if (!first) first = me;
    const size_t n = (size_t)std::min<uint64_t>(rest, per);
    rest -= n;
    put_be32(slot, rest ? (mNextPage == lock_page ? mNextPage + 1 : mNextPage) : 0);
    take(slot + 4, n);
    if (4 + n < mPageSize) memset(slot + 4 + n, 0, mPageSize - 4 - n);
  }
  return first;
}

void TDirectDbWriter::BuildPage(std::vector<unsigned char>& Page, uint32_t PageSize, unsigned char Type, const std::vector<std::string>& Cells, uint32_t RightMost)
{
  // PageSize here is the USABLE size; the caller's buffer already has the full page size
  const bool interior = (Type == 0x05 || Type == 0x02);
  const size_t hdr = interior ? 12 : 8;
  std::fill(Page.begin(), Page.end(), 0);
  size_t at = PageSize;
  for (size_t i = 0; i < Cells.size(); ++i) {
    at -= Cells[i].size();
    memcpy(Page.data() + at, Cells[i].data(), Cells[i].size());
    put_be16(Page.data() + hdr + 2 * i, (uint32_t)at);
  }
  if (hdr + 2 * Cells.size() > at) throw TReadableException("direct database load: page overflow");
  Page[0] = Type;
  put_be16(Page.data() + 3, (uint32_t)Cells.size());
  put_be16(Page.data() + 5, (uint32_t)(at == 65536 ? 0 : at));
  if (interior) put_be32(Page.data() + 8, RightMost);
}

void TDirectDbWriter::FlushLeaf(bool ToRoot)
{
  std::vector<unsigned char> page(mPageSize);
  BuildPage(page, mUsable, 0x0D, mLeafCells, 0);
  if (ToRoot) WritePageAt(mTableRoot, page.data());
  else {
    const uint32_t no = AppendPage(page.data());
    mTableChildren.push_back({ no, mLeafLastRow });
    mLeafFlushed = true;
  }
  mLeafCells.clear(); mLeafBytes = 0;
}

void TDirectDbWriter::AddRow(const std::vector<TDbValue>& Values)
{
  if (mFinished || Values.empty() || Values[0].mKind != TDbValue::kText) throw TReadableException("direct database load: a row needs its text key first");
  std::string key((const char*)Values[0].mData, Values[0].mSize);
  if (!mKeys.insert(key).second) throw TReadableException("direct database load: '" + key + "' is in the load twice");
  // record header + the values' bytes as segments (scalars are encoded into mScratch, texts / blobs are taken where they lie)
  std::string types;
  types.reserve(Values.size() * 2);
  mScratch.resize(Values.size() * 8);
  std::vector<Segment> segs;
  segs.reserve(Values.size() + 1);
  segs.push_back({ nullptr, 0 });                                   // the header, patched in below
  size_t sc = 0; uint64_t body = 0;
  auto add_seg = [&](const unsigned char* p, size_t n) {
    if (!n) return;
    if (segs.size() > 1 && segs.back().p + segs.back().n == p) segs.back().n += n; else segs.push_back({ p, n });
    body += n;
  };
  for (const TDbValue& v : Values) {
    switch (v.mKind) {
      case TDbValue::kNull: types.push_back((char)0); break;
      case TDbValue::kInt: { int nb; const int t = int_serial(v.mInt, mScratch.data() + sc, nb); types.push_back((char)t); add_seg(mScratch.data() + sc, (size_t)nb); sc += (size_t)nb; break; }
      case TDbValue::kReal: {
        uint64_t u; memcpy(&u, &v.mReal, 8); u = __builtin_bswap64(u);
        memcpy(mScratch.data() + sc, &u, 8); types.push_back((char)7); add_seg(mScratch.data() + sc, 8); sc += 8; break;
      }
      case TDbValue::kText: put_varint(types, 13 + 2 * (uint64_t)v.mSize); add_seg((const unsigned char*)v.mData, v.mSize); break;
      case TDbValue::kBlob: put_varint(types, 12 + 2 * (uint64_t)v.mSize); add_seg((const unsigned char*)v.mData, v.mSize); break;
    }
  }
  int hl = 1;
  while (varint_len(types.size() + (size_t)hl) != hl) ++hl;
  std::string header;
  put_varint(header, types.size() + (size_t)hl);
  header += types;
  segs[0] = { (const unsigned char*)header.data(), header.size() };
  const uint64_t payload = header.size() + body;

  const long long rowid = ++mRowId;
  std::string cell;
  put_varint(cell, payload);
  put_varint(cell, (uint64_t)rowid);
  const uint32_t local = LocalSize(payload, mUsable - 35);
  const size_t cell_size = cell.size() + local + (payload > local ? 4 : 0);
  if (8 + 2 * (mLeafCells.size() + 1) + mLeafBytes + cell_size > mUsable) FlushLeaf(false);
  const uint32_t first = SpillPayload(segs, payload, local, cell);
  if (first) cell += be32_string(first);
  mLeafBytes += cell.size();
  mLeafCells.push_back(std::move(cell));
  mLeafLastRow = rowid;
  mIndex.push_back({ std::move(key), rowid });
}

void TDirectDbWriter::FinishTable()
{
  if (mRowId == 0) return;
  if (!mLeafFlushed) { FlushLeaf(true); return; }                  // every cell fits the root leaf
  if (!mLeafCells.empty()) FlushLeaf(false);
  std::vector<std::pair<uint32_t, long long>> children = mTableChildren;
  const size_t cap = (mUsable - 12) / 15 + 1;                       // children per interior page: a cell is at most 4 + 9 bytes + its pointer
  std::vector<unsigned char> page(mPageSize);
  for (;;) {
    const size_t n = children.size();
    const size_t pages = (n + cap - 1) / cap;
    std::vector<std::pair<uint32_t, long long>> next;
    size_t at = 0;
    for (size_t g = 0; g < pages; ++g) {
      const size_t take = n / pages + (g < n % pages ? 1 : 0);      // even groups: every interior page gets at least two children
      std::vector<std::string> cells;
      for (size_t k = 0; k + 1 < take; ++k) { std::string c = be32_string(children[at + k].first); put_varint(c, (uint64_t)children[at + k].second); cells.push_back(std::move(c)); }
      BuildPage(page, mUsable, 0x05, cells, children[at + take - 1].first);
      if (pages == 1) { WritePageAt(mTableRoot, page.data()); return; }
      const uint32_t no = AppendPage(page.data());
      next.push_back({ no, children[at + take - 1].second });
      at += take;
    }
    children.swap(next);
  }
}

void TDirectDbWriter::FinishIndex()
{
  if (mIndex.empty()) return;
  std::sort(mIndex.begin(), mIndex.end());                          // BINARY collation = memcmp order, then the rowid
  struct Key { std::string cell; uint32_t left; };
  std::vector<Key> keys;
  keys.reserve(mIndex.size());
  const uint32_t max_local = ((mUsable - 12) * 64 / 255) - 23;
  for (const auto& e : mIndex) {
    unsigned char ib[8]; int nb;
    const int it = int_serial(e.second, ib, nb);
    std::string types;
    put_varint(types, 13 + 2 * (uint64_t)e.first.size());
    types.push_back((char)it);
    int hl = 1;
    while (varint_len(types.size() + (size_t)hl) != hl) ++hl;
    std::string rec;
    put_varint(rec, types.size() + (size_t)hl);
    rec += types; rec += e.first; rec.append((const char*)ib, (size_t)nb);
    Key k; k.left = 0;
    put_varint(k.cell, rec.size());
    const uint32_t local = LocalSize(rec.size(), max_local);
    std::vector<Segment> segs(1, Segment{ (const unsigned char*)rec.data(), rec.size() });
    const uint32_t first = SpillPayload(segs, rec.size(), local, k.cell);
    if (first) k.cell += be32_string(first);
    keys.push_back(std::move(k));
  }
  mIndex.clear(); mIndex.shrink_to_fit();
  bool leaf = true;
  uint32_t final_right = 0;
  for (;;) {
    // pack this level: a key that does not fit closes the page and moves up (index b-trees keep every key once)
    std::vector<std::vector<std::string>> pages(1);
    std::vector<uint32_t> rights;                                   // right-most child of every closed page (interior levels)
    std::vector<Key> up;
    size_t bytes = 0;
    const size_t hdr = leaf ? 8 : 12;
    for (size_t i = 0; i < keys.size(); ++i) {
      std::string cell = leaf ? keys[i].cell : be32_string(keys[i].left) + keys[i].cell;
      std::vector<std::string>& cur = pages.back();
      if (hdr + 2 * (cur.size() + 1) + bytes + cell.size() <= mUsable) { bytes += cell.size(); cur.push_back(std::move(cell)); continue; }
      if (i + 1 == keys.size()) {
        // the last key: the page after it must not be empty, so the key before it moves up instead
        if (cur.size() < 2) throw TReadableException("direct database load: index page too small");
        cur.pop_back();
        rights.push_back(leaf ? 0 : keys[i - 1].left);
        up.push_back(Key{ keys[i - 1].cell, 0 });
        pages.emplace_back();
        bytes = cell.size();
        pages.back().push_back(std::move(cell));
      } else {
        if (cur.empty()) throw TReadableException("direct database load: index page too small");
        rights.push_back(leaf ? 0 : keys[i].left);
        up.push_back(Key{ keys[i].cell, 0 });
        pages.emplace_back();
        bytes = 0;
      }
    }
    rights.push_back(leaf ? 0 : final_right);
    std::vector<unsigned char> page(mPageSize);
    if (pages.size() == 1) {
      BuildPage(page, mUsable, leaf ? 0x0A : 0x02, pages[0], rights[0]);
      WritePageAt(mIndexRoot, page.data());
      return;
    }
    for (size_t g = 0; g < pages.size(); ++g) {
      BuildPage(page, mUsable, leaf ? 0x0A : 0x02, pages[g], rights[g]);
      const uint32_t no = AppendPage(page.data());
      if (g < up.size()) up[g].left = no; else final_right = no;
    }
    keys.swap(up);
    leaf = false;
  }
}

void TDirectDbWriter::Finish()
{
  if (mFinished) return;
  mFinished = true;
  FinishTable();
  FinishIndex();
  FlushAppend();
  put_be32(mHeaderBuf.data() + 28, mNextPage - 1);                  // database size in pages
  const uint32_t change = get_be32(mHeaderBuf.data() + 24) + 1;
  put_be32(mHeaderBuf.data() + 24, change);                         // file change counter ...
  put_be32(mHeaderBuf.data() + 92, change);                         // ... and the "version valid for" stamp that makes the size above count
  full_pwrite(mFd, mHeaderBuf.data(), 100, 0);
  if (ftruncate(mFd, (off_t)(mNextPage - 1) * (off_t)mPageSize) != 0) throw TReadableException("direct database load: truncate failed");
  close(mFd); mFd = -1;
}

}  // namespace afec
