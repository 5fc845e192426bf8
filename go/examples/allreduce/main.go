// allreduce: the Go program a user of the reference writes once the collectives exist.
// Run with: gompirun 8 allreduce
// UNVERIFIED (no Go toolchain in the authoring image).
package main

import (
	"flag"
	"fmt"
	"log"
	"time"

	"github.com/btracey/mpi"
)

func main() {
	n := flag.Int("n", 1<<26, "float32 elements per rank (256 MiB)")
	flag.Parse()
	if err := mpi.Init(); err != nil {
		log.Fatal(err)
	}
	defer mpi.Finalize()

	send, err := mpi.Alloc([]float32(nil), *n) // device-resident: the zero-copy path
	if err != nil {
		log.Fatal(err)
	}
	recv, _ := mpi.Alloc([]float32(nil), *n)
	defer send.Free()
	defer recv.Free()

	const iters = 20
	mpi.Barrier()
	start := time.Now()
	for i := 0; i < iters; i++ {
		if err := mpi.Allreduce(send, recv, mpi.Sum); err != nil {
			log.Fatal(err)
		}
	}
	dt := time.Since(start).Seconds() / iters
	size := float64(mpi.Size())
	if mpi.Rank() == 0 {
		bytes := float64(*n) * 4
		fmt.Printf("allreduce %d MiB on %d ranks: %.3f ms, busbw %.1f GB/s\n",
			*n*4>>20, mpi.Size(), dt*1e3, bytes/dt*2*(size-1)/size/1e9)
	}

	// host slices work too (staged through the device heap): drop-in for existing callers
	x := []float64{float64(mpi.Rank()), 1}
	y := make([]float64, 2)
	if err := mpi.Allreduce(x, y, mpi.Sum); err != nil {
		log.Fatal(err)
	}
	fmt.Println("rank", mpi.Rank(), "sum of ranks and count:", y)
}
