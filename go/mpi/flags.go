package mpi

// flags.go -- the command-line contract of the reference (five -mpi-* flags, same names, same
// exported variables and flag.Value types: reference flags.go:10-50) plus the additive -mpi-gpu.
// UNVERIFIED (no Go toolchain in the authoring image).

import (
	"flag"
	"strings"
	"time"
)

// AddrsFlag collects comma separated addresses; repeated flags append.
type AddrsFlag []string

func (a *AddrsFlag) String() string { return strings.Join(*a, ",") }
func (a *AddrsFlag) Set(s string) error {
	*a = append(*a, strings.Split(s, ",")...)
	return nil
}

// DurationFlag is a time.Duration settable from the command line ("1.5s", "300ms").
type DurationFlag time.Duration

func (d *DurationFlag) String() string { return time.Duration(*d).String() }
func (d *DurationFlag) Set(s string) error {
	v, err := time.ParseDuration(s)
	if err == nil {
		*d = DurationFlag(v)
	}
	return err
}

// Values of the flags after flag.Parse(). Fields set on the Cuda struct take precedence.
var (
	FlagAddr        string       // -mpi-addr
	FlagAllAddrs    AddrsFlag    // -mpi-alladdr
	FlagInitTimeout DurationFlag // -mpi-inittimeout
	FlagProtocol    string       // -mpi-protocol (accepted; the control plane is always tcp)
	FlagPassword    string       // -mpi-password
	FlagGpu         int          // -mpi-gpu (new): CUDA ordinal, -1 = rank % device count
)

func init() {
	for _, f := range []struct {
		name, usage string
		str         *string
		def         string
	}{
		{"mpi-addr", "address of the local running process", &FlagAddr, ""},
		{"mpi-protocol", "communication protocol to use", &FlagProtocol, "tcp"},
		{"mpi-password", "value to use for salting the mpi connection", &FlagPassword, ""},
	} {
		flag.StringVar(f.str, f.name, f.def, f.usage)
	}
	flag.Var(&FlagAllAddrs, "mpi-alladdr", "addresses of all of the processes as comma separated values")
	flag.Var(&FlagInitTimeout, "mpi-inittimeout", "duration to wait before timeout in init")
	flag.IntVar(&FlagGpu, "mpi-gpu", -1, "CUDA device ordinal for this rank (-1: rank modulo device count)")
}
