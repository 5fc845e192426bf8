// Package mpi is the drop-in facade for github.com/btracey/mpi on a box of NVIDIA B200 GPUs.
//
// UNVERIFIED: this image has no Go toolchain, so these files have never been compiled.  They are
// the binding a maintainer adds; the C ABI they call (include/b200mpi.h) is exercised end to end
// by the C++ facade (mpi_b200/cpp/mpi.hpp) and by the Python harness through the same entry points.
//
// Callers keep writing
//
//	flag.Parse(); mpi.Init(); defer mpi.Finalize()
//	mpi.Send(x, dst, tag); mpi.Receive(&y, src, tag)
//
// exactly as with the reference (function set and signatures of reference mpi.go:96-159, the
// Interface of mpi.go:163-170, Register of mpi.go:61-67, Raw of mpi.go:75-91, TagExists of
// mpi.go:174-182).  What changes is the default implementation behind the facade: instead of
// Network (gob over TCP, network.go) it is Cuda (cuda.go), a cgo layer over libb200mpi.so whose
// kernels move the slices between GPUs through NVSwitch peer memory.
//
// New entry points, following the same conventions (package function forwarding to the
// registered implementation, blocking, error return, data typed by its Go slice type):
//
//	Recv       alias of Receive (the reference spells it Receive)
//	Bcast      root's slice to every rank
//	Allreduce  element-wise reduction over ranks, result on every rank
//	Allgather  concatenation in rank order on every rank
//	Barrier    rendezvous of all ranks
//	ReduceScatter, Reduce, Alltoall   the usual MPI companions of the above
//	Isend + Wait   the non-blocking Send design the reference sketches and comments out (mpi.go:132-152)
//
// They are served through the optional Collective interface, the upgrade the reference hints at
// with its unused isAllReducer variable (mpi.go:69-71).
package mpi

import "fmt"

// Interface is the set of routines an implementation must provide; unchanged from the reference.
type Interface interface {
	Init() error
	Finalize()
	Rank() int
	Size() int
	Send(data interface{}, destination, tag int) error
	Receive(data interface{}, source, tag int) error
}

// Op selects the reduction of Allreduce.
type Op int

const (
	Sum Op = iota
	Max
	Min
)

// Collective is implemented by transports that offer collectives natively.
type Collective interface {
	Bcast(data interface{}, root int) error
	Allreduce(send, recv interface{}, op Op) error
	Allgather(send, recv interface{}) error
	Barrier() error
}

// Collective2 is the second optional upgrade: further collectives and the Isend/Wait pair.
type Collective2 interface {
	ReduceScatter(send, recv interface{}, op Op) error
	Reduce(send, recv interface{}, op Op, root int) error
	Alltoall(send, recv interface{}) error
	Isend(data interface{}, destination, tag int) error
	Wait(destination, tag int) error
}

// Raw marks a payload that is sent as bytes without any encoding.
type Raw []byte

// GobEncode keeps Raw usable with gob-based implementations such as the reference's Network.
func (r Raw) GobEncode() ([]byte, error) { return append([]byte(nil), r...), nil }

// GobDecode is the inverse of GobEncode; it reuses the destination when it is large enough.
func (r *Raw) GobDecode(b []byte) error {
	if cap(*r) >= len(b) {
		*r = (*r)[:len(b)]
	} else {
		*r = make(Raw, len(b))
	}
	copy(*r, b)
	return nil
}

// TagExists reports that {peer, tag} already has a request in flight.
type TagExists struct {
	Tag int
}

func (t TagExists) Error() string { return fmt.Sprintf("Tag %v already in use sending", t.Tag) }

var (
	current    Interface = &Cuda{}
	registered bool
)

// Register installs an implementation; it may be called once, during program initialisation.
func Register(impl Interface) {
	if registered {
		panic("register called more than once")
	}
	current, registered = impl, true
}

// Init must precede every other call. flag.Parse() must have run if the -mpi-* flags are used.
func Init() error { return current.Init() }

// Finalize ends the session; no call may follow it.
func Finalize() { current.Finalize() }

// Rank is this process's index in the sorted address list, or -1 before Init.
func Rank() int { return current.Rank() }

// Size is the number of ranks, or 0 before Init.
func Size() int { return current.Size() }

// Send blocks until the matching Receive has taken the data. Concurrent calls need distinct
// {destination, tag} pairs.
func Send(data interface{}, destination, tag int) error {
	return current.Send(data, destination, tag)
}

// Receive blocks for the message {source, tag} and stores it through the pointer data, resizing
// the destination slice to the sent length.
func Receive(data interface{}, source, tag int) error {
	return current.Receive(data, source, tag)
}

// Recv is Receive.
func Recv(data interface{}, source, tag int) error { return Receive(data, source, tag) }

func collective() (Collective, error) {
	if c, ok := current.(Collective); ok {
		return c, nil
	}
	return nil, fmt.Errorf("mpi: registered implementation %T has no collectives", current)
}

// Bcast copies root's slice into data on every rank. data is a slice (or *DeviceSlice).
func Bcast(data interface{}, root int) error {
	c, err := collective()
	if err != nil {
		return err
	}
	return c.Bcast(data, root)
}

// Allreduce reduces send element-wise over all ranks into recv on every rank; send and recv are
// slices of the same type and length ([]float32, []float64, []int64) and may be the same slice.
func Allreduce(send, recv interface{}, op Op) error {
	c, err := collective()
	if err != nil {
		return err
	}
	return c.Allreduce(send, recv, op)
}

// Allgather stores rank r's send at recv[r*len(send):(r+1)*len(send)] on every rank.
func Allgather(send, recv interface{}) error {
	c, err := collective()
	if err != nil {
		return err
	}
	return c.Allgather(send, recv)
}

// Barrier returns once every rank has called it.
func Barrier() error {
	c, err := collective()
	if err != nil {
		return err
	}
	return c.Barrier()
}

func collective2() (Collective2, error) {
	if c, ok := current.(Collective2); ok {
		return c, nil
	}
	return nil, fmt.Errorf("mpi: registered implementation %T has no ReduceScatter/Reduce/Alltoall/Isend", current)
}

// ReduceScatter reduces block j of every rank's send (Size() blocks of len(recv)) into rank j's recv.
func ReduceScatter(send, recv interface{}, op Op) error {
	c, err := collective2()
	if err != nil {
		return err
	}
	return c.ReduceScatter(send, recv, op)
}

// Reduce is Allreduce with the result on root only; recv may be nil on the other ranks.
func Reduce(send, recv interface{}, op Op, root int) error {
	c, err := collective2()
	if err != nil {
		return err
	}
	return c.Reduce(send, recv, op, root)
}

// Alltoall sends block j of send to rank j, where it becomes block Rank() of recv.
func Alltoall(send, recv interface{}) error {
	c, err := collective2()
	if err != nil {
		return err
	}
	return c.Alltoall(send, recv)
}

// Isend transmits data to destination and returns once data may be modified again, without
// waiting for the receiver (the Send of the reference's commented-out design, mpi.go:132-143).
// The {destination, tag} pair stays in use until Wait.
func Isend(data interface{}, destination, tag int) error {
	c, err := collective2()
	if err != nil {
		return err
	}
	return c.Isend(data, destination, tag)
}

// Wait blocks until destination confirmed the message sent with tag, and frees the pair for
// re-use (mpi.go:146-152).
func Wait(destination, tag int) error {
	c, err := collective2()
	if err != nil {
		return err
	}
	return c.Wait(destination, tag)
}
