// main.go -- harness over the REAL github.com/emer/auditory sound.SndEnv, kept
// uncompiled here (no Go toolchain in the build image).  Run it wherever Go and
// the module cache exist to (1) re-pin the oracle -- it dumps MelFBankSegment,
// MFCCSegment, Energy and GborOutput of config 1 as raw little-endian float64 /
// float32 for tests to compare -- and (2) time the true Go CPU path.
//
//	go run ./baseline/go -wav sig.f32 -out dump.bin
package main

import (
	"encoding/binary"
	"flag"
	"fmt"
	"math"
	"os"
	"time"

	"github.com/emer/auditory/agabor"
	"github.com/emer/auditory/sound"
	"github.com/go-audio/audio"
)

func main() {
	in := flag.String("wav", "sig.f32", "raw little-endian float32 mono 16 kHz samples")
	out := flag.String("out", "dump.bin", "output dump")
	flag.Parse()
	raw, err := os.ReadFile(*in)
	if err != nil {
		panic(err)
	}
	n := len(raw) / 4
	se := &sound.SndEnv{}
	se.Defaults()
	se.Sound.Buf = &audio.IntBuffer{Format: &audio.Format{NumChannels: 1, SampleRate: 16000}}
	se.Signal.SetShape([]int{n}, nil, nil)
	for i := 0; i < n; i++ {
		se.Signal.Values[i] = float64(math.Float32frombits(binary.LittleEndian.Uint32(raw[4*i:])))
	}
	for _, or := range []float64{0, 45, 90, 135} {
		for _, ph := range []float64{0, 1.5708} {
			se.GaborSpecs = append(se.GaborSpecs, agabor.Filter{WaveLen: 2, Orientation: or, SigmaWidth: 0.5,
				SigmaLength: 0.5, PhaseOffset: ph, CircleEdge: true})
		}
	}
	se.GaborFilters.SizeX, se.GaborFilters.SizeY = 9, 9
	se.GaborFilters.StrideX, se.GaborFilters.StrideY = 3, 3
	se.GaborFilters.Gain = 2
	se.GborOutPoolsY, se.GborOutPoolsX, se.GborOutUnitsY, se.GborOutUnitsX = 8, 2, 2, 8
	if err := se.Init(); err != nil {
		panic(err)
	}
	f, _ := os.Create(*out)
	defer f.Close()
	t0 := time.Now()
	for seg := 0; seg < se.SegCnt; seg++ {
		se.ProcessSegment(seg, 0)
		g := se.ApplyGabor()
		binary.Write(f, binary.LittleEndian, se.MelFBankSegment.Values)
		binary.Write(f, binary.LittleEndian, se.MFCCSegment.Values)
		binary.Write(f, binary.LittleEndian, se.Energy.Values)
		binary.Write(f, binary.LittleEndian, g.Values)
	}
	dt := time.Since(t0).Seconds()
	fmt.Printf("segments %d  %.3f s  %.1f audio-s/s (1 goroutine)\n", se.SegCnt, dt, float64(n)/16000/dt)
}
