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
		{"mpi-addr", "this rank's own address (host:port or :port); its position in the sorted -mpi-alladdr list is the rank", &FlagAddr, ""},
		{"mpi-protocol", "kept for command-line compatibility; the control plane always speaks tcp", &FlagProtocol, "tcp"},
		{"mpi-password", "shared secret every rank must present during the Init handshake", &FlagPassword, ""},
	} {
		flag.StringVar(f.str, f.name, f.def, f.usage)
	}
	flag.Var(&FlagAllAddrs, "mpi-alladdr", "every rank's address, comma separated; may be given several times")
	flag.Var(&FlagInitTimeout, "mpi-inittimeout", "give up Init after this long (Go duration syntax); 0 waits forever")
	flag.IntVar(&FlagGpu, "mpi-gpu", -1, "CUDA device ordinal for this rank (-1: rank modulo device count)")
}
