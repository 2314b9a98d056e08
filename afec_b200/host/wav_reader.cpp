// RIFF / WAVE decoding on the host (the reference keeps decoding on the host too; north_star).
// Behaviour follows Source/Core/CoreFileFormats/Source/WaveFile.cpp:372-407 (chunk checks and error
// messages) and Export/SampleConverter.h:392-518 (sample value conventions: floats in 16-bit range).
#include "afx_host.h"

#include <cstdio>
#include <cstring>
#include <sys/stat.h>

namespace afec {

int ModificationStatTime(const std::string& FileName)
{
  struct stat st;
  return (stat(FileName.c_str(), &st) == 0) ? (int)st.st_mtime : 0;
}

std::string ExtractFileExtension(const std::string& FileName)
{
  const size_t slash = FileName.find_last_of('/');
  const size_t dot = FileName.find_last_of('.');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) return "";
  return FileName.substr(dot + 1);
}

static inline uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

void ReadWaveFile(const std::string& FileName, TDecodedAudio& Out)
{
  FILE* f = fopen(FileName.c_str(), "rb");
  if (!f) throw TReadableException("Failed to open the file for reading.");
  struct Closer { FILE* f; ~Closer() { fclose(f); } } closer{ f };
  struct stat st;
  Out.mFileSize = (fstat(fileno(f), &st) == 0) ? (int64_t)st.st_size : 0;

  unsigned char hdr[12];
  if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "RIFF", 4) != 0 || memcmp(hdr + 8, "WAVE", 4) != 0)
    throw TReadableException("Not a valid WAV file.");
  bool have_fmt = false, have_data = false;
  uint16_t tag = 0, channels = 0, bits = 0, block = 0; uint32_t rate = 0;
  long data_off = 0; uint32_t data_size = 0;
  long pos = 12;
  for (;;) {
    unsigned char ch[8];
    if (fseek(f, pos, SEEK_SET) != 0 || fread(ch, 1, 8, f) != 8) break;
    const uint32_t size = rd32(ch + 4);
    if (!memcmp(ch, "fmt ", 4) && !have_fmt) {
      unsigned char b[40] = { 0 };
      const size_t want = size < 40 ? size : 40;
      if (size < 16 || fread(b, 1, want, f) != want) throw TReadableException("Not a valid WAV file.");
      tag = rd16(b); channels = rd16(b + 2); rate = rd32(b + 4); block = rd16(b + 12); bits = rd16(b + 14);
      if (tag == 0xFFFE && size >= 26) tag = rd16(b + 24);      // WAVE_FORMAT_EXTENSIBLE: sub format
      have_fmt = true;
    } else if (!memcmp(ch, "data", 4) && !have_data) {
      data_off = pos + 8; data_size = size; have_data = true;
    }
    pos += 8 + (long)size + (size & 1);
    if (have_fmt && have_data) break;
  }
  if (!have_fmt || !have_data) throw TReadableException("Not a valid WAV file.");
  const bool pcm = (tag == 1), flt = (tag == 3);
  if ((!pcm && !flt) || channels == 0 || rate == 0 || rate > 0x7fffffffu || (pcm && bits != 8 && bits != 16 && bits != 24 && bits != 32) ||
      (flt && bits != 32 && bits != 64))
    throw TReadableException("Unsupported file format.");
  (void)block;
  const int bps = bits / 8;
  if ((int64_t)data_off + data_size > Out.mFileSize) data_size = (uint32_t)(Out.mFileSize > data_off ? Out.mFileSize - data_off : 0);
  const int64_t frames = (int64_t)data_size / ((int64_t)channels * bps);
  if (frames <= 0) throw TReadableException("Unsupported file format or corrupt file.");

  Out.mFrames = frames; Out.mChannels = channels; Out.mSampleRate = (int)rate; Out.mBitDepth = bits;
  const size_t n = (size_t)frames * channels;
  fseek(f, data_off, SEEK_SET);
  if (pcm && bits == 16) {                       // upload as is; the device converts (float)value
    Out.mFormat = AFX_PCM_I16;
    Out.mBytes.resize(n * 2);
    const size_t got = fread(Out.mBytes.data(), 1, n * 2, f);
    if (got < n * 2) memset(Out.mBytes.data() + got, 0, n * 2 - got);   // failed blocks are zeroed, SA.cpp:510-524
    return;
  }
  std::vector<unsigned char> raw(n * bps);
  const size_t got = fread(raw.data(), 1, raw.size(), f);
  if (got < raw.size()) memset(raw.data() + got, 0, raw.size() - got);
  Out.mFormat = AFX_PCM_F32;
  Out.mBytes.resize(n * 4);
  float* dst = reinterpret_cast<float*>(Out.mBytes.data());
  const unsigned char* p = raw.data();
  if (pcm && bits == 8) for (size_t i = 0; i < n; ++i) dst[i] = (float)(((int)p[i] - 128) << 8);
  else if (pcm && bits == 24) for (size_t i = 0; i < n; ++i, p += 3) {
    const int32_t v = (int32_t)(((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16)) << 8);
    dst[i] = (float)(v * 32768.0 / 2147483648.0);
  } else if (pcm && bits == 32) for (size_t i = 0; i < n; ++i, p += 4) {
    const double d = (double)(int32_t)rd32(p) * 32768.0 / 2147483648.0;
    const float v = (float)d; dst[i] = v < -32768.0f ? -32768.0f : (v > 32767.0f ? 32767.0f : v);
  } else if (flt && bits == 32) for (size_t i = 0; i < n; ++i, p += 4) {
    float v; memcpy(&v, p, 4);
    const double d = (double)v * 32768.0; dst[i] = (float)(d < -32768.0 ? -32768.0 : (d > 32767.0 ? 32767.0 : d));
  } else for (size_t i = 0; i < n; ++i, p += 8) {
    double v; memcpy(&v, p, 8);
    const double d = v * 32768.0; dst[i] = (float)(d < -32768.0 ? -32768.0 : (d > 32767.0 ? 32767.0 : d));
  }
}

}  // namespace afec
