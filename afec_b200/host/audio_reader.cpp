// RIFF / WAVE and AIFF / AIFC decoding on the host (the reference keeps decoding on the host too; north_star).
// "Decoding" here means finding the sample data: the bytes are handed to the GPU as they sit in the file
// (AFX_PCM_* raw formats, include/afec_b200.h) and converted there; only 64-bit float data is converted here.
//
// Behaviour follows Source/Core/CoreFileFormats/Source/WaveFile.cpp:372-407 and AifFile.cpp:150-372, 436-480 (chunk
// checks, accepted sample types, error messages) and Export/SampleConverter.h:392-518 (sample value conventions).
// ProbeAudioFile() reads headers only, ReadAudioData() reads the samples straight into caller memory -- a pinned ring
// slot in the batched extractor, so a file's bytes are touched once between the page cache and the H2D copy.
#include "afx_host.h"

#include <algorithm>
#include <cmath>
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

bool IsSupportedAudioFileExtension(const std::string& FileName)
{
  std::string e = ExtractFileExtension(FileName);
  std::transform(e.begin(), e.end(), e.begin(), ::tolower);
  return e == "wav" || e == "aif" || e == "aiff" || e == "aifc";
}

static inline uint32_t rd32(const unsigned char* p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t rd32be(const unsigned char* p) { return p[3] | (p[2] << 8) | (p[1] << 16) | ((uint32_t)p[0] << 24); }
static inline uint16_t rd16be(const unsigned char* p) { return (uint16_t)(p[1] | (p[0] << 8)); }

struct FileCloser { FILE* f; ~FileCloser() { if (f) fclose(f); } };

static void probe_wave(FILE* f, TAudioInfo& I)
{
  unsigned char hdr[12];
  if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "RIFF", 4) != 0 || memcmp(hdr + 8, "WAVE", 4) != 0)
    throw TReadableException("Not a valid WAV file.");
  bool have_fmt = false, have_data = false;
  uint16_t tag = 0, channels = 0, bits = 0; uint32_t rate = 0;
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
      tag = rd16(b); channels = rd16(b + 2); rate = rd32(b + 4); bits = rd16(b + 14);
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
  const int bps = bits / 8;
  if ((int64_t)data_off + data_size > I.mFileSize) data_size = (uint32_t)(I.mFileSize > data_off ? I.mFileSize - data_off : 0);
  const int64_t frames = (int64_t)data_size / ((int64_t)channels * bps);
  if (frames <= 0) throw TReadableException("Unsupported file format or corrupt file.");
  I.mFrames = frames; I.mChannels = channels; I.mSampleRate = (int)rate; I.mBitDepth = bits;
  I.mDataOffset = data_off; I.mFileBytesPerSample = bps; I.mHostConvert = 0;
  if (pcm) I.mFormat = bits == 8 ? AFX_PCM_U8 : bits == 16 ? AFX_PCM_I16 : bits == 24 ? AFX_PCM_I24 : AFX_PCM_I32;
  else if (bits == 32) I.mFormat = AFX_PCM_F32U;
  else { I.mFormat = AFX_PCM_F32; I.mHostConvert = 1; }          // 64-bit float, little endian
}

// 80-bit IEEE 754 extended (big endian), as the COMM chunk stores the sample rate
static double extended_to_double(const unsigned char* p)
{
  const int sign = p[0] >> 7, exp = ((p[0] & 0x7f) << 8) | p[1];
  uint64_t mant = 0;
  for (int i = 0; i < 8; ++i) mant = (mant << 8) | p[2 + i];
  if (exp == 0 && mant == 0) return 0.0;
  const double v = std::ldexp((double)mant, exp - 16383 - 63);
  return sign ? -v : v;
}

static void probe_aiff(FILE* f, TAudioInfo& I)
{
  unsigned char hdr[12];
  if (fread(hdr, 1, 12, f) != 12 || memcmp(hdr, "FORM", 4) != 0 || (memcmp(hdr + 8, "AIFF", 4) != 0 && memcmp(hdr + 8, "AIFC", 4) != 0))
    throw TReadableException("This is not a valid AIFF file!");
  const bool aifc = memcmp(hdr + 8, "AIFC", 4) == 0;
  bool have_comm = false, have_ssnd = false;
  unsigned char comm[64] = { 0 }; uint32_t comm_size = 0;
  long ssnd_off = 0;
  long pos = 12;
  for (;;) {
    unsigned char ch[8];
    if (fseek(f, pos, SEEK_SET) != 0 || fread(ch, 1, 8, f) != 8) break;
    const uint32_t size = rd32be(ch + 4);
    if (!memcmp(ch, "COMM", 4) && !have_comm) {
      comm_size = size < sizeof(comm) ? size : (uint32_t)sizeof(comm);
      if (size < 18 || fread(comm, 1, comm_size, f) != comm_size) throw TReadableException("This is not a valid AIFF file!");
      have_comm = true;
    } else if (!memcmp(ch, "SSND", 4) && !have_ssnd) { ssnd_off = pos + 8; have_ssnd = true; }
    pos += 8 + (long)size + (size & 1);
    if (have_comm && have_ssnd) break;
  }
  if (!have_comm || !have_ssnd) throw TReadableException("This is not a valid AIFF file!");
  const int channels = rd16be(comm), bits = rd16be(comm + 6);
  const uint32_t comm_frames = rd32be(comm + 2);
  const double rate = extended_to_double(comm + 8);
  if (bits != 8 && bits != 16 && bits != 24 && bits != 32 && bits != 64) throw TReadableException("Unsupported AIFF file type.");
  // AIFC compression types the reference accepts (AifFile.cpp:192-211, spelled there in its reversed four-cc convention):
  // big-endian PCM "NONE" / "twos", little-endian PCM "sowt", IEEE floats "fl32" / "fl64"
  bool little = false, is_float = false;
  if (aifc) {
    if (comm_size < 22) throw TReadableException("Unsupported compressed AIFC file type.");
    char ct[5] = { (char)comm[18], (char)comm[19], (char)comm[20], (char)comm[21], 0 };
    for (char* c = ct; *c; ++c) *c = (char)tolower(*c);
    if (!strcmp(ct, "none") || !strcmp(ct, "twos")) little = false;
    else if (!strcmp(ct, "sowt")) little = true;
    else if (!strcmp(ct, "fl32") || !strcmp(ct, "fl64")) is_float = true;
    else throw TReadableException("Unsupported compressed AIFC file type.");
  }
  if (channels <= 0 || !(rate >= 1.0) || rate > 2147483647.0) throw TReadableException("Unsupported AIFF file type.");
  unsigned char snd[8];
  if (fseek(f, ssnd_off, SEEK_SET) != 0 || fread(snd, 1, 8, f) != 8) throw TReadableException("This is not a valid AIFF file!");
  const int64_t data_off = (int64_t)ssnd_off + 8 + rd32be(snd);
  const int bps = bits / 8;
  const int64_t rest = (I.mFileSize > data_off) ? (I.mFileSize - data_off) / ((int64_t)channels * bps) : 0;
  const int64_t frames = std::min<int64_t>(rest, comm_frames);                  // AifFile.cpp:349-352
  I.mFrames = frames; I.mChannels = channels; I.mSampleRate = (int)rate; I.mBitDepth = bits;
  I.mDataOffset = data_off; I.mFileBytesPerSample = bps; I.mHostConvert = 0;
  if (bits == 64) { I.mFormat = AFX_PCM_F32; I.mHostConvert = 2; }             // 64-bit float, big endian
  else if (bits == 32 && is_float) I.mFormat = AFX_PCM_F32UBE;
  else if (bits == 8) I.mFormat = AFX_PCM_I8;
  else if (bits == 16) I.mFormat = little ? AFX_PCM_I16 : AFX_PCM_I16BE;
  else if (bits == 24) I.mFormat = little ? AFX_PCM_I24 : AFX_PCM_I24BE;
  else I.mFormat = little ? AFX_PCM_I32 : AFX_PCM_I32BE;
}

void ProbeAudioFile(const std::string& FileName, TAudioInfo& I)
{
  I = TAudioInfo();
  I.mFileName = FileName;
  FILE* f = fopen(FileName.c_str(), "rb");
  if (!f) throw TReadableException("Failed to open the file for reading.");
  FileCloser closer{ f };
  struct stat st;
  I.mFileSize = (fstat(fileno(f), &st) == 0) ? (int64_t)st.st_size : 0;
  std::string e = ExtractFileExtension(FileName);
  std::transform(e.begin(), e.end(), e.begin(), ::tolower);
  if (e == "aif" || e == "aiff" || e == "aifc") probe_aiff(f, I);
  else probe_wave(f, I);
  I.mDataBytes = (size_t)I.mFrames * (size_t)I.mChannels * (size_t)afx_pcm_bytes(I.mFormat);
}

void ReadAudioData(const TAudioInfo& I, unsigned char* Dst)
{
  if (I.mDataBytes == 0) return;
  FILE* f = fopen(I.mFileName.c_str(), "rb");
  if (!f) throw TReadableException("Failed to open the file for reading.");
  FileCloser closer{ f };
  fseek(f, (long)I.mDataOffset, SEEK_SET);
  const size_t n = (size_t)I.mFrames * (size_t)I.mChannels;
  if (!I.mHostConvert) {
    const size_t got = fread(Dst, 1, I.mDataBytes, f);
    if (got < I.mDataBytes) memset(Dst + got, 0, I.mDataBytes - got);    // failed blocks are zeroed, SA.cpp:510-524
    return;
  }
  // 64-bit floats: value * 32768, clamped, as float32 in 16-bit range
  std::vector<unsigned char> raw(n * 8);
  const size_t got = fread(raw.data(), 1, raw.size(), f);
  if (got < raw.size()) memset(raw.data() + got, 0, raw.size() - got);
  float* dst = reinterpret_cast<float*>(Dst);
  const unsigned char* p = raw.data();
  for (size_t i = 0; i < n; ++i, p += 8) {
    unsigned char b[8];
    if (I.mHostConvert == 2) for (int k = 0; k < 8; ++k) b[k] = p[7 - k]; else memcpy(b, p, 8);
    double v; memcpy(&v, b, 8);
    const double d = v * 32768.0;
    dst[i] = (float)(d < -32768.0 ? -32768.0 : (d > 32767.0 ? 32767.0 : d));
  }
}

void ReadAudioFile(const std::string& FileName, TDecodedAudio& Out)
{
  TAudioInfo I;
  ProbeAudioFile(FileName, I);
  Out.mFrames = I.mFrames; Out.mChannels = I.mChannels; Out.mSampleRate = I.mSampleRate; Out.mBitDepth = I.mBitDepth;
  Out.mFormat = I.mFormat; Out.mFileSize = I.mFileSize;
  Out.mBytes.resize(I.mDataBytes);
  ReadAudioData(I, Out.mBytes.data());
}

void ReadWaveFile(const std::string& FileName, TDecodedAudio& Out) { ReadAudioFile(FileName, Out); }

}  // namespace afec
