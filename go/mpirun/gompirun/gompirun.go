// gompirun launches one rank per local GPU.
//
//	gompirun N program [args...]      N may be "auto": one rank per GPU reported by nvidia-smi
//
// Same contract as the reference launcher (reference mpirun/gompirun/gompirun.go:28-93): ports
// ":6000"+i, user arguments first, then -mpi-addr / -mpi-alladdr, inherited stdio, wait for all.
// Added: -mpi-gpu <i mod ngpus> for every child and a non-zero exit status when a child fails.
// UNVERIFIED (no Go toolchain in the authoring image); mpirun/gompirun.cpp is the tested twin.
package main

import (
	"fmt"
	"os"
	"os/exec"
	"strconv"
	"strings"
	"sync"
)

const basePort = 6000

func gpuCount() int {
	if vis, ok := os.LookupEnv("CUDA_VISIBLE_DEVICES"); ok && vis != "" {
		return len(strings.Split(vis, ","))
	}
	out, err := exec.Command("nvidia-smi", "-L").Output()
	if err != nil {
		return 0
	}
	n := 0
	for _, line := range strings.Split(string(out), "\n") {
		if strings.HasPrefix(line, "GPU ") {
			n++
		}
	}
	return n
}

func main() {
	if len(os.Args) < 3 {
		fmt.Fprintln(os.Stderr, "usage: gompirun N program [args...]")
		os.Exit(2)
	}
	ngpu := gpuCount()
	n := ngpu
	if os.Args[1] != "auto" {
		v, err := strconv.Atoi(os.Args[1])
		if err != nil || v < 1 || v > 8 {
			fmt.Fprintln(os.Stderr, "gompirun: N must be 1..8 or auto")
			os.Exit(2)
		}
		n = v
	}
	if n < 1 {
		n = 1
	}
	addrs := make([]string, n)
	for i := range addrs {
		addrs[i] = ":" + strconv.Itoa(basePort+i)
	}
	list := strings.Join(addrs, ",")
	var wg sync.WaitGroup
	failed := make([]bool, n)
	for i := 0; i < n; i++ {
		args := append([]string{}, os.Args[3:]...)
		args = append(args, "-mpi-addr", addrs[i], "-mpi-alladdr", list)
		if ngpu > 0 {
			args = append(args, "-mpi-gpu", strconv.Itoa(i%ngpu))
		}
		cmd := exec.Command(os.Args[2], args...)
		cmd.Stdin, cmd.Stdout, cmd.Stderr = os.Stdin, os.Stdout, os.Stderr
		wg.Add(1)
		go func(i int) {
			defer wg.Done()
			failed[i] = cmd.Run() != nil
		}(i)
	}
	wg.Wait()
	for _, f := range failed {
		if f {
			os.Exit(1)
		}
	}
}
