// at3p_pipeline.cu — the body of TAt3PEnc::TImpl::EncodeFrame (src/atrac/at3p/at3p.cpp:88-194) for S streams
// x N consecutive lambda calls, on the device kernels of at3p_kernels.cu / at3p_gha.cu.
//
// Indexing.  Call t (t = 0 returns LOOK_AHEAD) filters PCM frame t into the PQF-domain frame P[t].  With
// A[n] = DoAnalize(cur = P[n], next = P[n+1]) (the GHA result of call n+1), call o+1 writes output frame o:
//     work  = P[o-1] (zeros for o = 0) minus the tones of A[o-1] ("now") and A[o] ("next"); the envelope
//             state the reference keeps in Atrac3pChanUnitCtx also needs A[o-2]            (at3p.cpp:133, ApplyFilter)
//     specs = MDCT(work / (32768 / 1.122018)), overlapping with the previous output's work  (:146-161)
//     frame = WriteFrame(tones = A[o-1], specs)                                             (:127-131, :168, :186-190)
// A batch continues t0 earlier calls: it carries P[t0-2], P[t0-1], the last 368 PCM samples (PQF history),
// the previous output's scaled residual (MDCT overlap), A[.] of the last two analyses and the
// ResultBufHistory envelopes.  Everything else is recomputed per batch; frames of a stream run in parallel.
#include "at3p_kernels.cuh"

#include <new>

namespace atde {
namespace at3p {

namespace {
template <class T> struct Buf {
    T* p = nullptr;
    size_t cap = 0;
    bool ensure(size_t n)
    {
        if (n <= cap) return true;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) return false;
        cap = n;
        return true;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
} // namespace

struct StreamState {
    int C = 0;
    int gha_flags = 7;                   // TAt3PEnc::TSettings::UseGha: 1 pass input, 2 write tonal, 4 write residual
    int n_streams = 0;                   // the carried arrays are sized (and zeroed) for this many streams
    long long calls = 0;                 // lambda calls every stream has seen
    long long pending = 0;               // calls of the batch being enqueued
    Buf<float> pcm_tail;                 // [S][368][C]
    Buf<float> band_hist;                // [S][C][2][2048]: P[t0-2], P[t0-1]
    Buf<float> resid_prev;               // [S][C][2048]
    Buf<ToneBlock> tone_hist;            // [S][2]: the last two analyses
    Buf<unsigned char> gha_hist;         // [S] GhaHistory
    struct Work {
        Buf<float> bands, resid, specs;
        Buf<ToneBlock> tones;
        Buf<unsigned char> frame_out, scratch;
    } w[2];
};

StreamState* pipeline_create(int C, int gha_flags)
{
    StreamState* st = new (std::nothrow) StreamState();
    if (st) { st->C = C; st->gha_flags = gha_flags; }
    return st;
}

void pipeline_destroy(StreamState* st)
{
    if (!st) return;
    st->pcm_tail.release(); st->band_hist.release(); st->resid_prev.release(); st->tone_hist.release(); st->gha_hist.release();
    for (auto& w : st->w) { w.bands.release(); w.resid.release(); w.specs.release(); w.tones.release(); w.frame_out.release(); w.scratch.release(); }
    delete st;
}

void pipeline_reset(StreamState* st)
{
    if (!st) return;
    st->n_streams = 0;
    st->calls = 0;
    st->pending = 0;
}

void pipeline_commit(StreamState* st)
{
    st->calls += st->pending;
    st->pending = 0;
}

#define PCK(call) do { if ((call) != cudaSuccess) { *err = #call; return -2; } } while (0)

int pipeline_run(StreamState* st, const float* d_pcm, int s0, int S, int total, long long N64, bool started,
                 unsigned char* d_out, cudaStream_t cs, int slot, long long* launches, const char** err,
                 const Profiler* prof)
{
    struct Scope {
        const Profiler* p; cudaStream_t s; int idx = -1;
        Scope(const Profiler* p_, cudaStream_t s_, int kind) : p(p_), s(s_) { if (p && p->begin) idx = p->begin(p->ctx, s, kind); }
        ~Scope() { if (p && p->end && idx >= 0) p->end(p->ctx, s, idx); }
    };
    const DevTables* T = device_tables();
    if (!T || !gha_tables_ready()) { *err = "ATRAC3plus table upload failed"; return -2; }
    const int C = st->C, N = (int)N64;
    if (st->n_streams != total) {
        // stream start: PqfCtx->buf, Buf1/Buf2/PrevBuf, MdctBuf zero-initialised (at3p.cpp:53-87, atrac3plus_pqf.c:104-119),
        // delay.NumToneBands = 0 (:47), no tones in ChUnit (at3p_gha.cpp:288-303), empty ResultBufHistory
        if (!st->pcm_tail.ensure((size_t)total * kPqfOverlap * C) || !st->band_hist.ensure((size_t)total * C * 2 * kFrame) ||
            !st->resid_prev.ensure((size_t)total * C * kFrame) || !st->tone_hist.ensure((size_t)total * 2) ||
            !st->gha_hist.ensure((size_t)total * gha_history_bytes())) { *err = "cudaMalloc (stream state)"; return -3; }
        PCK(cudaMemsetAsync(st->pcm_tail.p, 0, (size_t)total * kPqfOverlap * C * sizeof(float), cs));
        PCK(cudaMemsetAsync(st->band_hist.p, 0, (size_t)total * C * 2 * kFrame * sizeof(float), cs));
        PCK(cudaMemsetAsync(st->resid_prev.p, 0, (size_t)total * C * kFrame * sizeof(float), cs));
        PCK(cudaMemsetAsync(st->tone_hist.p, 0, (size_t)total * 2 * sizeof(ToneBlock), cs));
        PCK(cudaMemsetAsync(st->gha_hist.p, 0, (size_t)total * gha_history_bytes(), cs));
        PCK(cudaStreamSynchronize(cs));                       // the other slot's stream must see the zeros too
        st->n_streams = total;
        st->calls = 0;
    }
    (void)started;
    const long long t0 = st->calls;
    st->pending = N;
    const int L = N + 2;
    const int nA = (int)(t0 == 0 ? N - 1 : N);                // analyses == outputs of this batch
    const int jA0 = t0 == 0 ? 2 : 1;                          // band frame of the first analysis
    const int jW0 = t0 == 0 ? 1 : 0;                          // band frame the first output encodes
    const int TS = nA + 2;                                    // tone records per stream: two carried + new
    StreamState::Work& w = st->w[slot];
    const int gha_blocks = gha_blocks_for((long long)S * (nA > 0 ? nA : 1));
    // (sized for the continuation case, N analyses, also on the first batch with its N - 1: the second call of a run must
    // not re-allocate in mid-pipeline)
    const size_t nAmax = (size_t)(N > 0 ? N : 1);
    // the tone search's block count is NOT monotonic in the number of analyses (more analyses can mean larger batches per
    // block and fewer blocks): the scratch area covers this launch AND the continuation call's
    const int blocks_cont = gha_blocks_for((long long)S * (long long)nAmax);
    const int scratch_blocks = gha_blocks > blocks_cont ? gha_blocks : blocks_cont;
    if (!w.bands.ensure((size_t)S * C * L * kFrame) || !w.resid.ensure((size_t)S * C * (nAmax + 1) * kFrame) ||
        !w.specs.ensure((size_t)S * nAmax * C * kFrame) || !w.tones.ensure((size_t)S * (nAmax + 2)) ||
        !w.frame_out.ensure((size_t)S * nAmax * gha_frame_out_bytes()) ||
        !w.scratch.ensure(gha_scratch_bytes(scratch_blocks))) { *err = "cudaMalloc (workspace)"; return -3; }

    float* band_hist = st->band_hist.p + (size_t)s0 * C * 2 * kFrame;
    float* pcm_tail = st->pcm_tail.p + (size_t)s0 * kPqfOverlap * C;
    float* resid_prev = st->resid_prev.p + (size_t)s0 * C * kFrame;
    ToneBlock* tone_hist = st->tone_hist.p + (size_t)s0 * 2;
    unsigned char* gha_hist = st->gha_hist.p + (size_t)s0 * gha_history_bytes();

    // carried frames in front of the new ones
    PCK(cudaMemcpy2DAsync(w.bands.p, (size_t)L * kFrame * sizeof(float), band_hist, 2 * kFrame * sizeof(float),
                          2 * kFrame * sizeof(float), (size_t)S * C, cudaMemcpyDeviceToDevice, cs));
    PCK(cudaMemcpy2DAsync(w.tones.p, (size_t)TS * sizeof(ToneBlock), tone_hist, 2 * sizeof(ToneBlock),
                          2 * sizeof(ToneBlock), (size_t)S, cudaMemcpyDeviceToDevice, cs));
    PCK(cudaMemcpy2DAsync(w.resid.p, (size_t)(nA + 1) * kFrame * sizeof(float), resid_prev, kFrame * sizeof(float),
                          kFrame * sizeof(float), (size_t)S * C, cudaMemcpyDeviceToDevice, cs));
    { Scope sc(prof, cs, 0); launch_pqf(d_pcm, pcm_tail, w.bands.p, S, C, N, L, 2, cs); }
    *launches += 1;
    if (nA > 0) {
        {
            Scope sc(prof, cs, 3);
            launch_gha_search(w.bands.p, S, C, nA, L, jA0, w.scratch.p, w.frame_out.p, gha_blocks, cs);
            launch_gha_result(w.frame_out.p, S, C, nA, gha_hist, w.tones.p, TS, 2, cs);
        }
        FilterLayout lay;
        lay.fo = nA; lay.tone_stride = TS; lay.in_frames = L; lay.in_off = jW0; lay.out_frames = nA + 1; lay.out_off = 1;
        // the debug masks of `--advanced ghadbg=` (at3p.cpp:143-177): without PASS_INPUT the tones are subtracted from
        // zeros, without WRITE_RESIUDAL the MDCT sees zeros, without WRITE_TONAL no tone block is written
        const bool pass_input = st->gha_flags & 1, write_tonal = st->gha_flags & 2, write_resid = st->gha_flags & 4;
        { Scope sc(prof, cs, 4); launch_tone_filter(T, pass_input ? w.bands.p : nullptr, w.tones.p, w.tones.p + 1, w.tones.p + 2, w.resid.p, S * nA, C, lay, cs); }
        if (!write_resid) PCK(cudaMemsetAsync(w.resid.p, 0, (size_t)S * C * (nA + 1) * kFrame * sizeof(float), cs));
        { Scope sc(prof, cs, 0); launch_mdct(T, w.resid.p, w.specs.p, S, C, nA, 1, cs); }
        { Scope sc(prof, cs, 2); launch_pack(T, w.specs.p, write_tonal ? w.tones.p + 1 : nullptr, d_out, S * nA, C, nA, TS, cs); }
        *launches += 5;
    }
    // carry out
    PCK(cudaMemcpy2DAsync(band_hist, 2 * kFrame * sizeof(float), w.bands.p + (size_t)N * kFrame, (size_t)L * kFrame * sizeof(float),
                          2 * kFrame * sizeof(float), (size_t)S * C, cudaMemcpyDeviceToDevice, cs));
    PCK(cudaMemcpy2DAsync(pcm_tail, (size_t)kPqfOverlap * C * sizeof(float),
                          d_pcm + ((size_t)N * kFrame - kPqfOverlap) * C, (size_t)N * kFrame * C * sizeof(float),
                          (size_t)kPqfOverlap * C * sizeof(float), (size_t)S, cudaMemcpyDeviceToDevice, cs));
    if (nA > 0) {
        PCK(cudaMemcpy2DAsync(resid_prev, kFrame * sizeof(float), w.resid.p + (size_t)nA * kFrame, (size_t)(nA + 1) * kFrame * sizeof(float),
                              kFrame * sizeof(float), (size_t)S * C, cudaMemcpyDeviceToDevice, cs));
        PCK(cudaMemcpy2DAsync(tone_hist, 2 * sizeof(ToneBlock), w.tones.p + nA, (size_t)TS * sizeof(ToneBlock),
                              2 * sizeof(ToneBlock), (size_t)S, cudaMemcpyDeviceToDevice, cs));
    }
    if (cudaGetLastError() != cudaSuccess) { *err = "kernel launch"; return -2; }
    return 0;
}

} // namespace at3p
} // namespace atde
