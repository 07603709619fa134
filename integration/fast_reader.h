// fast_reader.h -- host-side reader for the libdbgb200 front end.
//
// The reference reads its inputs with std::getline over gzstream's 303-byte buffer on the main thread
// (DBG_contig/gzstream.h:47, DBGgraph.cpp:244-272), about 150 MB/s: once the build runs on the GPU that reader is
// what the wall clock shows.  This header keeps the reference's framing rules and replaces the mechanics:
//
//   LineSource    zlib gzread() (plain and .gz files alike) into a 4-MiB buffer, lines found with memchr, handed out
//                 as views -- no per-line std::string, no per-character stream calls;
//   FileProducer  one thread per input file frames reads (one-line FASTA / 4-line FASTQ, exactly the rules below)
//                 into page-locked blocks {bases, offsets}; a bounded queue hands the blocks to the main thread,
//                 which submits them to the GPU in file order -- so read ordinals, and with them the slot layout, are
//                 what a sequential reader would produce, while up to MAX_AHEAD files decode concurrently.
//
// Framing (DBGgraph.cpp:244-272): a line whose first byte is '@' (FASTQ, -f 1) or '>' (FASTA, -f 2) is a header;
// the NEXT line, whatever it contains, is the read (everything up to '\n', a '\r' included); FASTQ then skips two
// more lines.  Any other line is ignored.  A header that is the last line of the file yields an empty read.
#pragma once
#include <zlib.h>

#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace dbgio {

class LineSource {
public:
    explicit LineSource(const char *path, size_t buf_bytes = 4u << 20) : buf_(buf_bytes)
    {
        f_ = gzopen(path, "rb");
        if (f_) gzbuffer(f_, 1u << 20);
    }
    ~LineSource() { if (f_) gzclose(f_); }
    bool ok() const { return f_ != nullptr; }
    // non-empty after a read error (truncated / corrupt .gz, I/O error): the data delivered so far is a prefix
    const std::string &error() const { return err_; }

    // next line without its '\n'; the view stays valid until the next call.  false = no characters left
    // (the contract of std::getline on a good stream: a last line without '\n' is still delivered)
    bool next(const char *&p, size_t &n)
    {
        for (;;) {
            if (pos_ < end_) {
                const char *s = buf_.data() + pos_;
                const char *nl = static_cast<const char *>(memchr(s, '\n', end_ - pos_));
                if (nl) { p = s; n = (size_t)(nl - s); pos_ += n + 1; return true; }
            }
            if (eof_) {
                if (pos_ < end_) { p = buf_.data() + pos_; n = end_ - pos_; pos_ = end_; return true; }
                return false;
            }
            refill();
        }
    }

private:
    void refill()
    {
        // keep the unterminated tail, move it to the front, read more behind it (grow if one line fills the buffer)
        const size_t tail = end_ - pos_;
        if (tail && pos_) memmove(buf_.data(), buf_.data() + pos_, tail);
        pos_ = 0; end_ = tail;
        if (end_ == buf_.size()) buf_.resize(buf_.size() * 2);
        if (!f_) { eof_ = true; return; }
        const size_t want = buf_.size() - end_;
        const int got = gzread(f_, buf_.data() + end_, (unsigned)(want > (1u << 30) ? (1u << 30) : want));
        if (got <= 0) {
            eof_ = true;
            // gzread() <= 0 is a clean end of file only if zlib agrees (the reference's gzstream cannot tell: it would
            // silently use the prefix; here the front end gets to say so)
            int zerr = Z_OK;
            const char *msg = gzerror(f_, &zerr);
            if (got < 0 || (zerr != Z_OK && zerr != Z_STREAM_END)) err_ = msg && *msg ? msg : "read error";
        } else end_ += (size_t)got;
    }
    gzFile f_ = nullptr;
    std::string err_;
    std::vector<char> buf_;
    size_t pos_ = 0, end_ = 0;
    bool eof_ = false;
};

// one block of framed reads: reads i = bases[offs[i] .. offs[i+1])
struct ReadBlock {
    char *bases = nullptr;
    uint64_t *offs = nullptr;
    uint64_t cap_bases = 0, cap_reads = 0, n_reads = 0, n_bases = 0;
    bool last = false;      // the file's final block (possibly empty)
};

// memory hooks: the front end passes dbg_host_alloc / dbg_host_free so that blocks are page-locked
typedef int (*alloc_fn)(void **, uint64_t);
typedef int (*free_fn)(void *);

class FileProducer {
public:
    FileProducer(const std::string &path, int format, uint64_t trim_oversize, uint64_t block_bases, uint64_t block_reads,
                 alloc_fn a, free_fn f, int n_blocks = 2)
        : path_(path), format_(format), trim_(trim_oversize), cap_bases_(block_bases), cap_reads_(block_reads), alloc_(a), free_(f),
          n_blocks_(n_blocks)
    {
        th_ = std::thread([this]() { run(); });
    }
    ~FileProducer()
    {
        { std::lock_guard<std::mutex> g(m_); cancelled_ = true; }
        cv_.notify_all();
        if (th_.joinable()) th_.join();
        for (ReadBlock *b : all_) { if (b->bases) free_(b->bases); if (b->offs) free_(b->offs); delete b; }
    }
    // next filled block, in file order (blocks until one is ready); the caller gives it back with recycle()
    ReadBlock *pop()
    {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [this]() { return !ready_.empty(); });
        ReadBlock *b = ready_.front(); ready_.pop_front();
        return b;
    }
    void recycle(ReadBlock *b)
    {
        if (b == &sentinel_) return;
        { std::lock_guard<std::mutex> g(m_); b->n_reads = 0; b->n_bases = 0; b->last = false; free_list_.push_back(b); }
        cv_.notify_all();
    }
    bool failed() const { return failed_; }
    // input problem, if any, known once the file's last block has been popped: "" = none.  The reads delivered are
    // what the reference's reader would have used (nothing for a file that does not open, the decodable prefix of a
    // damaged .gz); the front end reports it instead of staying silent.
    std::string io_error()
    {
        std::lock_guard<std::mutex> g(m_);
        return io_error_;
    }

private:
    ReadBlock *get_free()
    {
        std::unique_lock<std::mutex> g(m_);
        if (free_list_.empty() && (int)all_.size() < n_blocks_) {
            g.unlock();
            ReadBlock *b = new ReadBlock();
            void *p = nullptr;
            if (alloc_(&p, cap_bases_) != 0) { delete b; failed_ = true; return nullptr; }
            b->bases = static_cast<char *>(p);
            if (alloc_(&p, (cap_reads_ + 1) * sizeof(uint64_t)) != 0) { free_(b->bases); delete b; failed_ = true; return nullptr; }
            b->offs = static_cast<uint64_t *>(p);
            b->cap_bases = cap_bases_; b->cap_reads = cap_reads_;
            g.lock();
            all_.push_back(b);
            return b;
        }
        cv_.wait(g, [this]() { return !free_list_.empty() || cancelled_; });
        if (cancelled_) return nullptr;
        ReadBlock *b = free_list_.front(); free_list_.pop_front();
        return b;
    }
    void publish(ReadBlock *b)
    {
        b->offs[b->n_reads] = b->n_bases;
        { std::lock_guard<std::mutex> g(m_); ready_.push_back(b); }
        cv_.notify_all();
    }
    void run()
    {
        LineSource src(path_.c_str());
        if (!src.ok()) { std::lock_guard<std::mutex> g(m_); io_error_ = "cannot open " + path_; }
        ReadBlock *b = get_free();
        if (!b) { publish_failure(); return; }
        const char hdr = format_ == 1 ? '@' : '>';
        const char *p; size_t n;
        while (src.ok() && src.next(p, n)) {
            if (n == 0 || p[0] != hdr) continue;
            const char *s = ""; size_t sn = 0;
            if (src.next(p, n)) { s = p; sn = n; }
            // a sequence larger than a whole block is cut to the -r length: only the first maxReadLen bases of a read
            // are ever used (DBGgraph.cpp:63); the only effect is on the logged, untrimmed k-mer count
            if (sn > cap_bases_) sn = trim_;
            if (b->n_reads == b->cap_reads || b->n_bases + sn > b->cap_bases) {
                publish(b);
                b = get_free();
                if (!b) { publish_failure(); return; }
            }
            b->offs[b->n_reads++] = b->n_bases;
            memcpy(b->bases + b->n_bases, s, sn);     // before the skips below: they invalidate the view
            b->n_bases += sn;
            if (format_ == 1) { const char *q; size_t qn; if (src.next(q, qn)) src.next(q, qn); }
        }
        if (!src.error().empty()) { std::lock_guard<std::mutex> g(m_); io_error_ = path_ + ": " + src.error() + " (reads up to the damage were used)"; }
        b->last = true;
        publish(b);
    }
    void publish_failure()
    {
        failed_ = true;
        sentinel_.last = true;          // empty, last: wakes the consumer, which checks failed()
        { std::lock_guard<std::mutex> g(m_); ready_.push_back(&sentinel_); }
        cv_.notify_all();
    }

    std::string path_;
    int format_;
    uint64_t trim_, cap_bases_, cap_reads_;
    alloc_fn alloc_; free_fn free_;
    int n_blocks_;
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<ReadBlock *> ready_, free_list_;
    std::vector<ReadBlock *> all_;
    ReadBlock sentinel_;
    std::string io_error_;
    bool cancelled_ = false;
    volatile bool failed_ = false;
};

}   // namespace dbgio
