// dvb_stream.cu — the reference's shared-memory example stream (--stream_examples / fast_pipeline), both ends (host code).
//
// Wire format and hand-shake restated from
//   deepvariant/stream_examples.cc:60-92      producer attach (open_only; items_available and shard_finished taken at start)
//   deepvariant/stream_examples.cc:94-156     WriteLenToShm / WriteBytesToShm / StreamExample
//   deepvariant/stream_examples.cc:158-176    StartStreaming / EndStreaming / SignalShardFinished
//   deepvariant/stream_examples_kernel.cc:166-240   consumer: try items_available, drain the buffer, release buffer_empty
//   deepvariant/fast_pipeline.cc:125-165      orchestrator: create / remove the objects;  fast_pipeline_utils.h:44-64 their names
//
// One buffer per make_examples shard:  { int32 len, alt_allele_indices bytes, int32 len, variant bytes, int32 len, image bytes } ...
// closed by int32 0 (native-endian ints, as the reference memcpy's them).  Three named mutexes per shard order the two processes:
//   buffer_empty     held by whoever owns the buffer's write side;  items_available  released by the producer when a buffer is ready;
//   shard_finished   released by the producer when its shard is done.
// boost::interprocess::shared_memory_object / named_mutex are, on Linux, POSIX shm_open("/name") and a named semaphore with initial
// count 1 (lock = sem_wait, try_lock = sem_trywait, unlock = sem_post) - that is what this file uses, so either end can be the
// reference's own binary.

#include <fcntl.h>
#include <semaphore.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "dvb_common.h"

struct DvbStream {
  std::string prefix;
  int shard = 0;
  int role = 0;                  // DVB_STREAM_ORCHESTRATOR / PRODUCER / CONSUMER
  int64_t size = 0;
  unsigned char* buf = nullptr;
  int64_t pos = 0;
  sem_t* buffer_empty = SEM_FAILED;
  sem_t* items_available = SEM_FAILED;
  sem_t* shard_finished = SEM_FAILED;
  bool finished = false;         // consumer: this shard's shard_finished was observed
  // consumer: the records of the last drained buffer
  std::vector<uint8_t> alt_blob, variant_blob;
  std::vector<int64_t> alt_begin, variant_begin;
};

namespace {

std::string ObjName(const std::string& prefix, const char* what, int shard) { return "/" + prefix + what + std::to_string(shard); }

int OpenAll(DvbStream* s, bool create) {
  const std::string shm = ObjName(s->prefix, "_shm_", s->shard);
  const int fd = shm_open(shm.c_str(), create ? (O_CREAT | O_RDWR) : O_RDWR, 0644);
  if (fd < 0) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "shm_open(%s): %s", shm.c_str(), strerror(errno));
  if (create && ftruncate(fd, s->size) != 0) { close(fd); return dvb::fail(DVB_ERR_INTERNAL, "ftruncate(%s): %s", shm.c_str(), strerror(errno)); }
  struct stat st;
  if (fstat(fd, &st) != 0 || st.st_size < 16) { close(fd); return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "%s: bad size", shm.c_str()); }
  s->size = st.st_size;
  void* p = mmap(nullptr, (size_t)s->size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return dvb::fail(DVB_ERR_INTERNAL, "mmap(%s): %s", shm.c_str(), strerror(errno));
  s->buf = static_cast<unsigned char*>(p);
  auto sem = [&](const char* what, sem_t** out) -> int {
    const std::string n = ObjName(s->prefix, what, s->shard);
    *out = create ? sem_open(n.c_str(), O_CREAT, 0644, 1) : sem_open(n.c_str(), 0);
    if (*out == SEM_FAILED) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "sem_open(%s): %s", n.c_str(), strerror(errno));
    return DVB_OK;
  };
  int st2;
  if ((st2 = sem("_buffer_empty_", &s->buffer_empty))) return st2;
  if ((st2 = sem("_items_available_", &s->items_available))) return st2;
  if ((st2 = sem("_shard_finished_", &s->shard_finished))) return st2;
  return DVB_OK;
}

void CloseAll(DvbStream* s) {
  if (s->buf) munmap(s->buf, (size_t)s->size);
  if (s->buffer_empty != SEM_FAILED) sem_close(s->buffer_empty);
  if (s->items_available != SEM_FAILED) sem_close(s->items_available);
  if (s->shard_finished != SEM_FAILED) sem_close(s->shard_finished);
}

// named_mutex::lock: a signal that interrupts the wait must not be taken for the lock
inline void Lock(sem_t* m) { while (sem_wait(m) != 0 && errno == EINTR) {} }

inline void WriteLen(DvbStream* s, int len) { memcpy(s->buf + s->pos, &len, sizeof len); s->pos += sizeof len; }
inline void WriteBytes(DvbStream* s, const void* p, int len) { WriteLen(s, len); memcpy(s->buf + s->pos, p, (size_t)len); s->pos += len; }

}  // namespace

extern "C" {

int dvb_stream_open(const char* shm_prefix, int32_t shard, int32_t role, int64_t buffer_size, DvbStream** out) {
  if (!shm_prefix || !out || shard < 0 || role < 0 || role > 2 || (role == DVB_STREAM_ORCHESTRATOR && buffer_size < 16))
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_stream_open: bad arguments");
  *out = nullptr;
  DvbStream* s = new DvbStream();
  s->prefix = shm_prefix; s->shard = shard; s->role = role; s->size = buffer_size;
  const int st = OpenAll(s, role == DVB_STREAM_ORCHESTRATOR);
  if (st) { CloseAll(s); delete s; return st; }
  if (role == DVB_STREAM_PRODUCER) {   // stream_examples.cc:83-91: nothing is available and the shard is not finished yet
    Lock(s->items_available);
    Lock(s->shard_finished);
  }
  *out = s;
  return DVB_OK;
}

void dvb_stream_close(DvbStream* s) {
  if (!s) return;
  CloseAll(s);
  delete s;
}

// fast_pipeline.cc:155-165: the orchestrator removes the objects when the run is over.
int dvb_stream_remove(const char* shm_prefix, int32_t shard) {
  if (!shm_prefix) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_stream_remove: null prefix");
  const std::string p = shm_prefix;
  shm_unlink(ObjName(p, "_shm_", shard).c_str());
  sem_unlink(ObjName(p, "_buffer_empty_", shard).c_str());
  sem_unlink(ObjName(p, "_items_available_", shard).c_str());
  sem_unlink(ObjName(p, "_shard_finished_", shard).c_str());
  return DVB_OK;
}

int64_t dvb_stream_buffer_size(const DvbStream* s) { return s ? s->size : 0; }

// ---- producer (make_examples side) ---------------------------------------------------------------------------------------
int dvb_stream_start(DvbStream* s) {   // StartStreaming: once per region
  if (!s || s->role != DVB_STREAM_PRODUCER) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_stream_start: not a producer");
  Lock(s->buffer_empty);
  s->pos = 0;
  return DVB_OK;
}

int dvb_stream_put(DvbStream* s, const void* alt_indices, int32_t alt_len, const void* variant, int32_t variant_len, const uint8_t* image,
                   int32_t image_len) {   // StreamExample
  if (!s || s->role != DVB_STREAM_PRODUCER || alt_len < 0 || variant_len < 0 || image_len < 0)
    return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_stream_put: bad arguments");
  const int64_t required = (int64_t)variant_len + alt_len + image_len + (int64_t)sizeof(int) * 4;
  if (required >= s->size) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "one example (%lld bytes) does not fit the %lld-byte buffer", (long long)required, (long long)s->size);
  while (true) {
    if (required < s->size - s->pos) {
      WriteBytes(s, alt_indices, alt_len);
      WriteBytes(s, variant, variant_len);
      WriteLen(s, image_len);
      memcpy(s->buf + s->pos, image, (size_t)image_len);
      s->pos += image_len;
      return DVB_OK;
    }
    if (s->size - s->pos > 0) WriteLen(s, 0);   // end of this buffer's batch (the reference writes it when any byte is left)
    sem_post(s->items_available);
    Lock(s->buffer_empty);
    s->pos = 0;
  }
}

int dvb_stream_end(DvbStream* s, int32_t data_written) {   // EndStreaming: once per region
  if (!s || s->role != DVB_STREAM_PRODUCER) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_stream_end: not a producer");
  if (data_written) {
    WriteLen(s, 0);
    sem_post(s->items_available);
  } else {
    sem_post(s->buffer_empty);
  }
  return DVB_OK;
}

int dvb_stream_shard_finished(DvbStream* s) {   // SignalShardFinished
  if (!s || s->role != DVB_STREAM_PRODUCER) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_stream_shard_finished: not a producer");
  Lock(s->buffer_empty);
  sem_post(s->shard_finished);
  return DVB_OK;
}

// The reference starts call_variants beside its make_examples processes and relies on the model load being slower than the producers'
// attach (a consumer that polls first would take the still-unlocked items_available / shard_finished mutexes).  A consumer here can
// wait for the attach instead: the producer of a shard is attached once it holds items_available and shard_finished (both 0), or holds
// buffer_empty (a region is open or the shard already ended).  Returns DVB_OK, or DVB_ERR_INTERNAL after timeout_ms.
int dvb_stream_wait_attached(DvbStream* s, int64_t timeout_ms) {
  if (!s) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_stream_wait_attached: null stream");
  for (int64_t waited = 0;; waited += 1) {
    int be = 1, ia = 1, sf = 1;
    sem_getvalue(s->buffer_empty, &be); sem_getvalue(s->items_available, &ia); sem_getvalue(s->shard_finished, &sf);
    if ((ia == 0 && sf == 0) || be == 0) return DVB_OK;
    if (waited >= timeout_ms) return dvb::fail(DVB_ERR_INTERNAL, "no producer attached to stream shard %d within %lld ms", s->shard, (long long)timeout_ms);
    usleep(1000);
  }
}

// ---- consumer (call_variants side) ---------------------------------------------------------------------------------------
// StreamExamplesResource::Next over `n` shards, starting at shard `index % n`: the first shard with a ready buffer is drained - its
// images are copied to images_host (capacity images_cap bytes; every image must have image_bytes bytes), its variant / alt-index
// records stay readable through *meta until the next call on that shard.  *n_out = 0 and *all_finished = 1 when every shard has
// signalled completion; *n_out = 0 and *all_finished = 0 when the shard that was polled last has just finished (call again).
// Busy-polls like the reference (try_lock round-robin), yielding between rounds.
int dvb_stream_next(DvbStream* const* shards, int32_t n, int64_t index, uint8_t* images_host, int64_t images_cap, int64_t image_bytes,
                    int32_t* n_out, int32_t* shard_out, DvbExampleBatchMeta* meta, int32_t* all_finished) {
  if (!shards || n < 1 || !n_out || !all_finished || image_bytes < 0) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_stream_next: bad arguments");
  *n_out = 0;
  *all_finished = 0;
  if (shard_out) *shard_out = -1;
  int shard = (int)(index % n);
  while (true) {
    int done = 0;
    for (int i = 0; i < n; ++i) done += shards[i]->finished ? 1 : 0;
    if (done >= n) { *all_finished = 1; return DVB_OK; }
    DvbStream* s = shards[shard];
    if (s->role != DVB_STREAM_CONSUMER) return dvb::fail(DVB_ERR_INVALID_ARGUMENT, "dvb_stream_next: shard %d is not a consumer handle", shard);
    if (!s->finished && sem_trywait(s->items_available) == 0) {
      s->alt_blob.clear(); s->variant_blob.clear();
      s->alt_begin.assign(1, 0); s->variant_begin.assign(1, 0);
      int64_t pos = 0, used = 0;
      int count = 0, len = 0;
      auto read_len = [&]() {       // a length that would lie past the buffer reads as -1 (malformed: every branch below rejects it)
        if (pos + (int64_t)sizeof len > s->size) { len = -1; return; }
        memcpy(&len, s->buf + pos, sizeof len); pos += sizeof len;
      };
      int st = DVB_OK;
      read_len();
      if (len < 0) st = dvb::fail(DVB_ERR_INTERNAL, "stream shard %d: bad record length", shard);
      while (len > 0) {
        if (pos + len > s->size) { st = dvb::fail(DVB_ERR_INTERNAL, "stream shard %d: record runs past the buffer", shard); break; }
        s->alt_blob.insert(s->alt_blob.end(), s->buf + pos, s->buf + pos + len); pos += len;
        s->alt_begin.push_back((int64_t)s->alt_blob.size());
        read_len();
        if (len < 0 || pos + len > s->size) { st = dvb::fail(DVB_ERR_INTERNAL, "stream shard %d: bad variant length", shard); break; }
        s->variant_blob.insert(s->variant_blob.end(), s->buf + pos, s->buf + pos + len); pos += len;
        s->variant_begin.push_back((int64_t)s->variant_blob.size());
        read_len();
        if (len != image_bytes || pos + len > s->size) { st = dvb::fail(DVB_ERR_INVALID_ARGUMENT, "stream shard %d: image of %d bytes, expected %lld", shard, len, (long long)image_bytes); break; }
        if (used + len > images_cap) { st = dvb::fail(DVB_ERR_INVALID_ARGUMENT, "stream shard %d: the buffer holds more images than images_host can take", shard); break; }
        memcpy(images_host + used, s->buf + pos, (size_t)len); pos += len; used += len;
        ++count;
        if (pos + (int64_t)sizeof len > s->size) break;   // a completely full buffer carries no terminator
        read_len();
      }
      sem_post(s->buffer_empty);
      if (st) return st;
      *n_out = count;
      if (shard_out) *shard_out = shard;
      if (meta) {
        meta->alt_blob = s->alt_blob.data(); meta->alt_begin = s->alt_begin.data();
        meta->variant_blob = s->variant_blob.data(); meta->variant_begin = s->variant_begin.data();
      }
      return DVB_OK;
    }
    if (!s->finished && sem_trywait(s->shard_finished) == 0) {
      s->finished = true;
      if (shard_out) *shard_out = shard;
      return DVB_OK;            // like the reference: an empty batch for this call; the caller comes back
    }
    shard = (shard + 1) % n;
    if (shard == (int)(index % n)) usleep(200);
  }
}

}  // extern "C"
