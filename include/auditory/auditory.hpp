// auditory.hpp -- C++ host-side mirror of the emer/auditory Go API for the
// speech-feature path, over the C-ABI in auditory_b200.h.  Header only.
//
// The reference is compiled Go and no Go toolchain exists in the build image,
// so this is the compiled-language drop-in: same type and method names as the
// Go packages (sound.SndEnv, sound.Params, dft.Params, mel.Params,
// mel.FilterBank, agabor.Filter, agabor.FilterSet), same argument meaning and
// the same error behaviour (Init reports errors, Process* leave the tensors
// untouched on failure).  Tensors are etensor-like: a shape and one contiguous
// row-major float vector (float32: SURVEY F4).  File:line citations refer to
// the reference tree.
#ifndef AUDITORY_AUDITORY_HPP_
#define AUDITORY_AUDITORY_HPP_

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../auditory_b200.h"

namespace auditory {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
inline void check(int32_t rc) {
    if (rc < 0) throw Error(rc, aud_last_error());
}

namespace etensor {
struct Float32 {   // emer/etable etensor.Float32: Shape + Values
    std::vector<int> Shp;
    std::vector<float> Values;
    void SetShape(std::vector<int> shp) {
        Shp = std::move(shp);
        size_t n = 1;
        for (int d : Shp) n *= (size_t)d;
        Values.assign(n, 0.f);
    }
    int Dim(int i) const { return Shp[i]; }
    int NumDims() const { return (int)Shp.size(); }
    size_t Len() const { return Values.size(); }
    void SetZeros() { std::fill(Values.begin(), Values.end(), 0.f); }
    float FloatValRowCell(int row, int cell) const { return Values[(size_t)row * (Len() / Shp[0]) + cell]; }
};
struct Float64 {
    std::vector<int> Shp;
    std::vector<double> Values;
    void SetShape(std::vector<int> shp) {
        Shp = std::move(shp);
        size_t n = 1;
        for (int d : Shp) n *= (size_t)d;
        Values.assign(n, 0.0);
    }
};
}  // namespace etensor

namespace dft {
struct Params {   // dft/dft.go:15-31
    bool CompLogPow = true;
    double LogMin = -100, LogOffSet = 1.0, PrevSmooth = 0, CurSmooth = 1.0;
    void Defaults() {   // dft/dft.go:33-39
        PrevSmooth = 0;
        CurSmooth = 1.0 - PrevSmooth;
        CompLogPow = true;
        LogOffSet = 1.0;
        LogMin = -100;
    }
};
}  // namespace dft

namespace mel {
inline double FreqToMel(double f) { return aud_freq_to_mel(f); }
inline double MelToFreq(double m) { return aud_mel_to_freq(m); }
inline int FreqToBin(double f, double nfft, double sr) { return aud_freq_to_bin(f, nfft, sr); }

struct FilterBank {   // mel/mel.go:16-44
    int NFilters = 32;
    double LoHz = 0, HiHz = 8000, LogOff = 0, LogMin = -10;
    bool Renorm = true;
    double RenormMin = -6, RenormMax = 4, RenormScale = 0;
    void Defaults() {   // mel/mel.go:171-180
        LoHz = 0; HiHz = 8000; NFilters = 32; LogOff = 0; LogMin = -10; Renorm = true; RenormMin = -6; RenormMax = 4;
    }
};
struct Params {   // mel/mel.go:47-66
    FilterBank FBank;
    std::vector<int32_t> BinPts;
    std::vector<double> HzPts;
    bool MFCC = false, Deltas = false;
    int NCoefs = 13;
    void Defaults() {   // mel/mel.go:69-74 (MFCC and Deltas ON)
        FBank.Defaults();
        MFCC = true;
        NCoefs = 13;
        Deltas = true;
    }
    void InitFilters(int dftSize, int sampleRate, etensor::Float64 *filters) {   // mel/mel.go:77-117
        BinPts.assign(FBank.NFilters + 2, 0);
        HzPts.assign(FBank.NFilters + 2, 0.0);
        FBank.Renorm = false;
        filters->SetShape({FBank.NFilters, FBank.NFilters + 2});
        check(aud_mel_init_filters(dftSize, sampleRate, FBank.NFilters, FBank.LoHz, FBank.HiHz, BinPts.data(),
                                   HzPts.data(), filters->Values.data()));
    }
};
}  // namespace mel

namespace agabor {
struct Filter {   // agabor/gabor.go:17-42
    bool Off = false;
    double WaveLen = 0, Orientation = 0, SigmaWidth = 0, SigmaLength = 0, PhaseOffset = 0;
    bool CircleEdge = false, Circular = false;
};
struct FilterSet {   // agabor/gabor.go:45-70
    int SizeX = 0, SizeY = 0, StrideX = 0, StrideY = 0;
    double Gain = 0;
    bool Distribute = false;
    etensor::Float64 Filters;
};
inline std::vector<Filter> Active(const std::vector<Filter> &specs) {   // agabor/gabor.go:329-336
    std::vector<Filter> a;
    for (const auto &s : specs)
        if (!s.Off) a.push_back(s);
    return a;
}
inline void ToTensor(const std::vector<Filter> &specs, FilterSet *set) {   // agabor/gabor.go:89-222
    std::vector<aud_gabor_spec> cs;
    for (const auto &s : specs)
        cs.push_back({s.Off ? 1 : 0, s.WaveLen, s.Orientation, s.SigmaWidth, s.SigmaLength, s.PhaseOffset,
                      s.CircleEdge ? 1 : 0, s.Circular ? 1 : 0});
    const int n = (int)Active(specs).size();
    set->Filters.SetShape({n, set->SizeY, set->SizeX});
    if (n) check(aud_gabor_to_tensor(cs.data(), (int)cs.size(), set->SizeX, set->SizeY, set->Distribute ? 1 : 0,
                                     set->Filters.Values.data()));
}
// agabor/gabor.go:225-315 on the GPU.  melData: [NFilters, steps] (or n of them back to back); rawOut: a 2-D
// or 4-D tensor (or n of them), written in place; cells the convolution does not reach keep their values.
inline void Convolve(const etensor::Float32 &melData, const FilterSet &filters, etensor::Float32 *rawOut, bool byTime,
                     int n = 1, int device = 0) {
    const int nd = melData.NumDims();
    const int nMel = melData.Shp[nd - 2], steps = melData.Shp[nd - 1];
    const int od = rawOut->NumDims() - (n > 1 ? 1 : 0);
    int32_t shape[4] = {0, 0, 0, 0};
    for (int i = 0; i < od && i < 4; ++i) shape[i] = rawOut->Shp[i + (n > 1 ? 1 : 0)];
    check(aud_gabor_convolve(device, melData.Values.data(), n, nMel, steps, filters.Filters.Values.data(),
                             filters.Filters.Shp.empty() ? 0 : filters.Filters.Shp[0], filters.SizeX, filters.SizeY,
                             filters.StrideX, filters.StrideY, filters.Gain, od, shape, byTime ? 1 : 0,
                             rawOut->Values.data()));
}
}  // namespace agabor

namespace sound {
inline int MSecToSamples(double ms, int rate) { return aud_msec_to_samples(ms, rate); }   // sndenv.go:522-524
inline double SamplesToMSec(int samples, int rate) { return 1000.0 * samples / rate; }

// sound.Wave (sound/sound.go:32-141): a decoded WAV file.  Load stands in for go-audio/wav's
// Decoder.FullPCMBuffer (third-party, unpinned): integer PCM, little endian, 8 (unsigned, as go-audio
// reads it) / 16 / 24 / 32 bits, channels interleaved in Data.
struct Wave {
    std::vector<int32_t> Data;   // audio.IntBuffer.Data
    int SourceBitDepth = 16, NumChannels = 1, Rate = 0;

    bool Load(const std::string &fn, std::string *err = nullptr) {
        auto failed = [&](const char *m) { if (err) *err = std::string("sound.Load: ") + m; return false; };
        std::FILE *f = std::fopen(fn.c_str(), "rb");
        if (!f) return failed("couldn't open file");
        std::vector<unsigned char> raw;
        unsigned char buf[1 << 16];
        size_t got;
        while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) raw.insert(raw.end(), buf, buf + got);
        std::fclose(f);
        if (raw.size() < 12 || std::memcmp(raw.data(), "RIFF", 4) || std::memcmp(raw.data() + 8, "WAVE", 4))
            return failed("not a RIFF/WAVE file");
        auto u32 = [&](size_t i) { return (uint32_t)raw[i] | (uint32_t)raw[i + 1] << 8 | (uint32_t)raw[i + 2] << 16 | (uint32_t)raw[i + 3] << 24; };
        auto u16 = [&](size_t i) { return (uint32_t)raw[i] | (uint32_t)raw[i + 1] << 8; };
        size_t pos = 12, fmt = 0, data = 0, dlen = 0;
        while (pos + 8 <= raw.size()) {
            const size_t size = u32(pos + 4);
            if (!std::memcmp(raw.data() + pos, "fmt ", 4)) fmt = pos + 8;
            else if (!std::memcmp(raw.data() + pos, "data", 4)) { data = pos + 8; dlen = std::min(size, raw.size() - data); break; }
            pos += 8 + size + (size & 1);
        }
        if (!fmt || !data || fmt + 16 > raw.size()) return failed("no fmt / data chunk");
        const uint32_t tag = u16(fmt);
        if (tag != 1 && tag != 0xFFFE) return failed("only integer PCM is supported");
        NumChannels = (int)u16(fmt + 2);
        Rate = (int)u32(fmt + 4);
        SourceBitDepth = (int)u16(fmt + 14);
        const int bytes = SourceBitDepth / 8;
        if (bytes < 1 || bytes > 4 || SourceBitDepth % 8) return failed("unsupported bit depth");
        Data.assign(dlen / bytes, 0);
        for (size_t i = 0; i < Data.size(); ++i) {
            const unsigned char *q = raw.data() + data + i * bytes;
            switch (bytes) {
                case 1: Data[i] = q[0]; break;
                case 2: Data[i] = (int16_t)(q[0] | q[1] << 8); break;
                case 3: { int32_t v = q[0] | q[1] << 8 | q[2] << 16; Data[i] = (v & 0x800000) ? v - (1 << 24) : v; break; }
                default: Data[i] = (int32_t)((uint32_t)q[0] | (uint32_t)q[1] << 8 | (uint32_t)q[2] << 16 | (uint32_t)q[3] << 24);
            }
        }
        return true;
    }
    int SampleRate() const { return Rate; }
    int Channels() const { return NumChannels; }
    int NumFrames() const { return NumChannels > 0 ? (int)(Data.size() / NumChannels) : 0; }
    double GetFloatAtIdx(size_t idx) const {   // sound/sound.go:130-141
        switch (SourceBitDepth) {
            case 32: return (double)Data[idx] / (double)0x7FFFFFFF;
            case 24: return (double)Data[idx] / (double)0x7FFFFF;
            case 16: return (double)Data[idx] / (double)0x7FFF;
            case 8: return (double)Data[idx] / (double)0x7F;
        }
        return 0;
    }
    // sound/sound.go:116-127: Data[i] for i < NumFrames (for interleaved multi-channel data: the first
    // NumFrames interleaved samples, as in the reference), narrowed to the float32 the GPU path consumes
    bool SoundToTensor(etensor::Float32 *samples) const {
        const int n = NumFrames();
        samples->SetShape({n});
        for (int i = 0; i < n; ++i) samples->Values[i] = (float)GetFloatAtIdx((size_t)i);
        return true;
    }
};

struct Params {   // sound/sndenv.go:24-61
    double WinMs = 25, StepMs = 10, SegmentMs = 100, StrideMs = 100;
    int BorderSteps = 2, Channel = 0;
    int WinSamples = 0, StepSamples = 0, SegmentSamples = 0, StrideSamples = 0, SegmentSteps = 0;
    std::vector<int> Steps;
};

// All per-segment outputs of a batch, [total segments][...] row-major.
struct BatchOutputs {
    std::vector<float> Mel, MFCC, Deltas, DeltaDeltas, Energy, Gabor;
    int64_t Segments = 0;
};

namespace kwta {
// vision kwta.KWTA / kwta.NeighInhib (third party, see auditory_b200.h): the parameter block of aud_apply_kwta with the
// reference's field meanings; Defaults() = KWTA.Defaults().
struct KWTA : aud_kwta_params {
    KWTA() { aud_kwta_defaults(this); on = 0; }
    bool On() const { return on != 0; }
    void Defaults() { const int pm = pool_mode, no = neigh_on; const float ng = neigh_gi; aud_kwta_defaults(this); pool_mode = pm; neigh_on = no; neigh_gi = ng; }
};
struct NeighInhib {
    bool On = false;
    float Gi = 0.6f;
    void Defaults() { On = true; Gi = 0.6f; }
};
}  // namespace kwta

class SndEnv {   // sound/sndenv.go:73-182
  public:
    std::string Nm, Dsc;
    bool On = true;
    Params Params_;
    Params &P() { return Params_; }
    Wave Sound;
    etensor::Float32 Signal;
    int SampleRate = 0, Channels = 1;   // Sound.SampleRate() / Sound.Channels() (set by ToTensor, or directly)
    int SegCnt = 0;
    dft::Params DFT;
    mel::Params Mel;
    etensor::Float64 MelFilters;
    etensor::Float32 MelFBankSegment, Energy, MFCCSegment, MFCCDeltas, MFCCDeltaDeltas, GborOutput, GborKwta, ExtGi;
    kwta::KWTA Kwta;
    kwta::NeighInhib NeighInhib;
    bool KwtaPool = false;
    std::vector<agabor::Filter> GaborSpecs;
    agabor::FilterSet GaborFilters;
    int GborOutPoolsX = 0, GborOutPoolsY = 0, GborOutUnitsX = 0, GborOutUnitsY = 0;
    bool ByTime = false;
    int Device = 0;

    ~SndEnv() { Close(); }
    void Close() {
        if (handle_) aud_destroy(handle_);
        handle_ = nullptr;
    }

    void ParamDefaults() {   // sndenv.go:64-71
        Params_.WinMs = 25; Params_.StepMs = 10; Params_.SegmentMs = 100; Params_.Channel = 0;
        Params_.StrideMs = 100; Params_.BorderSteps = 2;
    }
    void Defaults() {   // sndenv.go:185-192
        ParamDefaults();
        On = true;
        Mel.Defaults();
        Kwta.Defaults();      // kwta ON and pool mode, as in the reference (sndenv.go:189-190)
        KwtaPool = true;
        ByTime = false;
    }

    // sndenv.go:195-267.  Returns false (with the message in err) where the Go code returns an error.
    bool Init(std::string *err = nullptr) {
        if (SampleRate <= 0) {
            if (err) *err = "sample rate <= 0";
            return false;
        }
        Params &p = Params_;
        p.WinSamples = MSecToSamples(p.WinMs, SampleRate);
        p.StepSamples = MSecToSamples(p.StepMs, SampleRate);
        p.SegmentSamples = MSecToSamples(p.SegmentMs, SampleRate);
        p.SegmentSteps = (int)std::round(p.SegmentMs / p.StepMs) + 2 * p.BorderSteps;
        p.StrideSamples = MSecToSamples(p.StrideMs, SampleRate);
        agabor::ToTensor(agabor::Active(GaborSpecs), &GaborFilters);
        nGabor_ = (int)agabor::Active(GaborSpecs).size();
        if (GborOutPoolsX == 0 && GborOutPoolsY == 0) GborOutput.SetShape({GborOutUnitsY, GborOutUnitsX});
        else if (GborOutPoolsX > 0 && GborOutPoolsY > 0)
            GborOutput.SetShape({GborOutPoolsY, GborOutPoolsX, GborOutUnitsY, GborOutUnitsX});
        else {
            if (err) *err = "GborOutPoolsX & GborOutPoolsY must both be == 0 or > 0 (i.e. 2D or 4D)";
            return false;
        }
        DFT.Defaults();   // wipes smoothing set before Init
        Mel.InitFilters(p.WinSamples, SampleRate, &MelFilters);
        p.Steps.resize(p.SegmentSteps);
        for (int i = 0; i < p.SegmentSteps; ++i) p.Steps[i] = p.StepSamples * (i - p.BorderSteps);
        MelFBankSegment.SetShape({Mel.FBank.NFilters, p.SegmentSteps});
        Energy.SetShape({p.SegmentSteps});
        if (Mel.MFCC) {
            MFCCSegment.SetShape({Mel.NCoefs, p.SegmentSteps});
            MFCCDeltas.SetShape({Mel.NCoefs, p.SegmentSteps});
            MFCCDeltaDeltas.SetShape({Mel.NCoefs, p.SegmentSteps});
        }
        int siglen = (int)Signal.Values.size() - p.SegmentSamples * Channels;
        siglen = siglen / Channels;
        SegCnt = siglen / p.StrideSamples + 1;
        Close();
        cacheValid_ = false;
        return true;
    }

    // Every segment of every utterance of a batch in one fused launch.
    BatchOutputs ProcessBatch(const float *wave, const std::vector<int64_t> &uttOffset, const std::vector<int32_t> &uttLen,
                              int addMs = 0) {
        ensureHandle();
        BatchOutputs out;
        out.Segments = aud_total_segments(handle_, uttLen.data(), (int)uttLen.size(), nullptr);
        aud_dims d{};
        check(aud_get_dims(handle_, &d));
        const size_t S = d.segment_steps, n = (size_t)out.Segments;
        out.Mel.assign(n * d.n_mel * S, 0.f);
        out.Energy.assign(n * S, 0.f);
        aud_outputs o{};
        o.mel = out.Mel.data();
        o.energy = out.Energy.data();
        if (Mel.MFCC) {
            out.MFCC.assign(n * d.n_coefs * S, 0.f);
            o.mfcc = out.MFCC.data();
            if (Mel.Deltas) {
                out.Deltas.assign(n * d.n_coefs * S, 0.f);
                out.DeltaDeltas.assign(n * d.n_coefs * S, 0.f);
                o.deltas = out.Deltas.data();
                o.delta_deltas = out.DeltaDeltas.data();
            }
        }
        if (nGabor_) {
            out.Gabor.assign(n * (size_t)d.gabor_len, 0.f);
            o.gabor = out.Gabor.data();
        }
        aud_batch b{wave, uttOffset.data(), uttLen.data(), (int32_t)uttLen.size(), MSecToSamples(addMs, SampleRate)};
        check(aud_process_host(handle_, &b, &o));
        return out;
    }

    // sndenv.go:342-433, one segment per call as in the reference; the GPU work for the whole signal is
    // done by the first call for a given `add` and later calls are served from that result.
    void ProcessSegment(int segment, int add) {
        if (!cacheValid_ || cacheAdd_ != add) {
            cache_ = ProcessBatch(Signal.Values.data(), {0}, {(int32_t)Signal.Values.size()}, add);
            cacheAdd_ = add;
            cacheValid_ = true;
            kwtaValid_ = false;
        }
        if (segment < 0 || segment >= cache_.Segments) throw std::out_of_range("segment");
        copyOut(cache_.Mel, MelFBankSegment, segment);
        copyOut(cache_.Energy, Energy, segment);
        if (Mel.MFCC) {
            copyOut(cache_.MFCC, MFCCSegment, segment);
            if (Mel.Deltas) {
                copyOut(cache_.Deltas, MFCCDeltas, segment);
                copyOut(cache_.DeltaDeltas, MFCCDeltaDeltas, segment);
            }
        }
        segment_ = segment;
    }
    // sndenv.go:481-497: GborOutput of the segment last processed, then ApplyNeighInhib / ApplyKwta (:303-323).  The kwta
    // step runs once for all segments of the signal, in segment order (KWTAPool keeps per-pool state from call to call).
    etensor::Float32 *ApplyGabor() {
        if (!(nGabor_ && cacheValid_)) return &GborOutput;
        copyOut(cache_.Gabor, GborOutput, segment_);
        if (!Kwta.On() && !NeighInhib.On) return &GborOutput;
        if (!kwtaValid_) {
            aud_kwta_params kp = Kwta;
            kp.pool_mode = KwtaPool ? 1 : 0;
            kp.neigh_on = NeighInhib.On ? 1 : 0;
            kp.neigh_gi = NeighInhib.Gi;
            std::vector<int32_t> shp(GborOutput.Shp.begin(), GborOutput.Shp.end());
            kwta_.assign(cache_.Gabor.size(), 0.f);
            extGi_.assign(cache_.Gabor.size(), 0.f);
            if (aud_apply_kwta(Device, &kp, cache_.Gabor.data(), (int32_t)cache_.Segments, (int32_t)shp.size(), shp.data(), nullptr, 0,
                               extGi_.data(), kwta_.data()) != AUD_OK)
                throw std::runtime_error(aud_last_error());
            kwtaValid_ = true;
        }
        ExtGi.SetShape(GborOutput.Shp);
        GborKwta.SetShape(GborOutput.Shp);
        copyOut(extGi_, ExtGi, segment_);
        copyOut(kwta_, GborKwta, segment_);
        return Kwta.On() ? &GborKwta : &GborOutput;
    }
    int Tail(size_t signalLen) const {   // sndenv.go:503-507
        return (int)(((long)signalLen - Params_.SegmentSamples) % Params_.StrideSamples);
    }
    // sndenv.go:510-519: pad so that the length of the signal divided by the stride has no remainder
    std::vector<float> Pad(const std::vector<float> &signal, float value) const {
        const int tail = Tail(signal.size());
        const int padLen = Params_.SegmentSamples - Params_.StepSamples - tail % Params_.StepSamples;
        std::vector<float> padded(signal);
        padded.insert(padded.end(), (size_t)std::max(0, padLen), value);
        return padded;
    }
    // sndenv.go:297-300
    bool ToTensor() {
        Sound.SoundToTensor(&Signal);
        SampleRate = Sound.SampleRate();
        Channels = Sound.Channels();
        cacheValid_ = false;
        return true;
    }
    // sndenv.go:274-294: trim or prepend leading silence (milliseconds); returns the offset
    int AdjustForSilence(double add, double existing) {
        if (SampleRate <= 0) return -1;
        int offset = 0;
        if (add >= 0) {
            if (add < existing) {
                offset = (int)(existing - add);
                const size_t n = std::min(Signal.Values.size(), (size_t)MSecToSamples((double)offset, SampleRate));
                Signal.Values.erase(Signal.Values.begin(), Signal.Values.begin() + n);
            } else if (add > existing) {
                offset = (int)(add - existing);
                Signal.Values.insert(Signal.Values.begin(), (size_t)MSecToSamples((double)offset, SampleRate), 0.f);
            }
            Signal.Shp = {(int)Signal.Values.size()};
        }
        cacheValid_ = false;
        return offset;
    }

  private:
    aud_handle *handle_ = nullptr;
    int nGabor_ = 0, segment_ = 0, cacheAdd_ = 0;
    bool cacheValid_ = false, kwtaValid_ = false;
    BatchOutputs cache_;
    std::vector<float> kwta_, extGi_;   // GborKwta / ExtGi of every segment of the cached signal

    static void copyOut(const std::vector<float> &src, etensor::Float32 &dst, int64_t seg) {
        std::memcpy(dst.Values.data(), src.data() + (size_t)seg * dst.Values.size(), dst.Values.size() * sizeof(float));
    }
    void ensureHandle() {
        if (handle_) return;
        const Params &p = Params_;
        aud_params ap{};
        ap.sample_rate = SampleRate;
        ap.win_samples = p.WinSamples; ap.step_samples = p.StepSamples; ap.segment_samples = p.SegmentSamples;
        ap.stride_samples = p.StrideSamples; ap.segment_steps = p.SegmentSteps; ap.border_steps = p.BorderSteps;
        ap.comp_log_pow = DFT.CompLogPow; ap.log_min = DFT.LogMin; ap.log_offset = DFT.LogOffSet;
        ap.prev_smooth = DFT.PrevSmooth; ap.cur_smooth = DFT.CurSmooth;
        ap.n_mel = Mel.FBank.NFilters; ap.mel_log_off = Mel.FBank.LogOff; ap.mel_log_min = Mel.FBank.LogMin;
        ap.renorm = Mel.FBank.Renorm; ap.renorm_min = Mel.FBank.RenormMin; ap.renorm_scale = Mel.FBank.RenormScale;
        ap.mfcc = Mel.MFCC; ap.n_coefs = Mel.NCoefs; ap.deltas = Mel.MFCC && Mel.Deltas; ap.mfcc_c0_energy = 1;
        ap.gabor_nf = nGabor_;
        ap.gabor_size_x = GaborFilters.SizeX; ap.gabor_size_y = GaborFilters.SizeY;
        ap.gabor_stride_x = GaborFilters.StrideX; ap.gabor_stride_y = GaborFilters.StrideY;
        ap.gabor_gain = GaborFilters.Gain;
        ap.gabor_out_dims = GborOutput.NumDims();
        for (int i = 0; i < GborOutput.NumDims(); ++i) ap.gabor_shape[i] = GborOutput.Shp[i];
        ap.gabor_by_time = ByTime;
        check(aud_create(&ap, Mel.BinPts.data(), MelFilters.Values.data(),
                         nGabor_ ? GaborFilters.Filters.Values.data() : nullptr, nullptr, Device, &handle_));
    }
};
}  // namespace sound
}  // namespace auditory
#endif  // AUDITORY_AUDITORY_HPP_
