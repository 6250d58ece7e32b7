// sndenv_demo.cpp -- the reference's call sequence through the C++ mirror:
//   se.Defaults(); fill Signal; set gabor specs; se.Init(); for seg: ProcessSegment(seg, 0); ApplyGabor()
// Prints one line per requested segment with checksums, compared by tests/test_cpp_mirror.py with the
// Python mirror (same library underneath) and the oracle.
#include <cstdio>
#include <cstdlib>

#include "auditory/auditory.hpp"

int main(int argc, char **argv) {
    using namespace auditory;
    const int n = argc > 1 ? std::atoi(argv[1]) : 32000;
    sound::SndEnv se;
    se.Defaults();
    if (argc > 2 && std::string(argv[2]) == "wav") {   // sndenv_demo 0 wav <file>: Sound.Load + ToTensor + Init
        std::string werr;
        if (argc < 4 || !se.Sound.Load(argv[3], &werr)) {
            std::fprintf(stderr, "%s\n", werr.c_str());
            return 4;
        }
        se.ToTensor();
        std::string ierr;
        if (!se.Init(&ierr)) {
            std::fprintf(stderr, "Init: %s\n", ierr.c_str());
            return 2;
        }
        double sum = 0;
        for (float v : se.Signal.Values) sum += v;
        std::printf("rate %d channels %d frames %d bits %d SegCnt %d WinSamples %d sum %.9f\n", se.Sound.SampleRate(),
                    se.Sound.Channels(), se.Sound.NumFrames(), se.Sound.SourceBitDepth, se.SegCnt, se.P().WinSamples, sum);
        return 0;
    }
    se.SampleRate = 16000;
    se.Signal.SetShape({n});
    unsigned s = 12345u;   // deterministic LCG noise + tone (no <random> distribution differences across libstdc++)
    for (int i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        se.Signal.Values[i] = 0.1f * ((float)(s >> 8) / 8388608.0f - 1.0f) + 0.3f * std::sin(2.0 * M_PI * 1000.0 * i / 16000.0);
    }
    for (double orient : {0.0, 45.0, 90.0, 135.0})
        for (double ph : {0.0, 1.5708}) {
            agabor::Filter f;
            f.WaveLen = 2; f.Orientation = orient; f.SigmaWidth = 0.5; f.SigmaLength = 0.5; f.PhaseOffset = ph; f.CircleEdge = true;
            se.GaborSpecs.push_back(f);
        }
    se.GaborFilters.SizeX = se.GaborFilters.SizeY = 9;
    se.GaborFilters.StrideX = se.GaborFilters.StrideY = 3;
    se.GaborFilters.Gain = 2;
    se.GborOutPoolsY = 8; se.GborOutPoolsX = 2; se.GborOutUnitsY = 2; se.GborOutUnitsX = 8;
    std::string err;
    if (!se.Init(&err)) {
        std::fprintf(stderr, "Init: %s\n", err.c_str());
        return 2;
    }
    std::printf("SegCnt %d SegmentSteps %d WinSamples %d\n", se.SegCnt, se.P().SegmentSteps, se.P().WinSamples);
    if (argc > 2 && std::string(argv[2]) == "init-only") return 0;
    try {
        for (int seg : {0, se.SegCnt / 2, se.SegCnt - 1}) {
            se.ProcessSegment(seg, 0);
            const etensor::Float32 *g = se.ApplyGabor();
            double sm = 0, sc = 0, sg = 0;
            for (float v : se.MelFBankSegment.Values) sm += v;
            for (float v : se.MFCCSegment.Values) sc += v;
            for (float v : g->Values) sg += v;
            std::printf("seg %d mel %.6f mfcc %.6f gabor %.6f e0 %.6f\n", seg, sm, sc, sg, se.Energy.Values[0]);
        }
    } catch (const Error &e) {
        std::fprintf(stderr, "error %d: %s\n", e.code, e.what());
        return 3;
    }
    return 0;
}
