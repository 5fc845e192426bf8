package mpi

// cuda.go -- the cgo shim: one Go method per C entry point of include/b200mpi.h.
// UNVERIFIED (no Go toolchain in the authoring image); see the package comment in mpi.go.
//
// Build: CGO_CFLAGS="-I<repo>/include" CGO_LDFLAGS="-L<repo>/mpi_b200/lib -lb200mpi" go build
//
// cgo rules observed here: Go pointers are handed to C only for the duration of one call and
// are never retained by the library (it stages or finishes before returning); empty slices pass
// a nil pointer (count 0 is legal in the ABI; bounce sends length-0 messages); blocking C calls
// occupy an OS thread, which is fine because every mpi call is blocking by contract.

/*
#cgo LDFLAGS: -lb200mpi
#include <stdlib.h>
#include "b200mpi.h"
*/
import "C"

import (
	"bytes"
	"encoding/gob"
	"errors"
	"fmt"
	"strings"
	"time"
	"unsafe"
)

// Cuda implements Interface and Collective on libb200mpi. Like Network in the reference, zero
// fields are filled from the -mpi-* flags at Init.
type Cuda struct {
	Addr     string        // address of the local process (-mpi-addr)
	Addrs    []string      // addresses of all processes (-mpi-alladdr)
	Timeout  time.Duration // Init fails after this long (-mpi-inittimeout); 0 waits forever
	Password string        // compared during the handshake (-mpi-password)
	Gpu      *int          // CUDA ordinal; nil takes -mpi-gpu (default -1: rank % device count)
}

const (
	dtU8  = C.B200MPI_U8
	dtI64 = C.B200MPI_I64
	dtF32 = C.B200MPI_F32
	dtF64 = C.B200MPI_F64
)

// DeviceSlice is device memory from the rank's peer-mapped heap; collectives and Send/Receive
// use it in place (no host staging).
type DeviceSlice struct {
	Ptr   unsafe.Pointer
	Len   int
	dtype C.int
}

// Alloc returns a device slice of n elements shaped like the zero-length prototype, e.g.
// mpi.Alloc([]float32(nil), 1<<26).
func Alloc(prototype interface{}, n int) (*DeviceSlice, error) {
	var dt C.int
	var es int
	switch prototype.(type) {
	case []float32:
		dt, es = dtF32, 4
	case []float64:
		dt, es = dtF64, 8
	case []int64:
		dt, es = dtI64, 8
	case []byte:
		dt, es = dtU8, 1
	default:
		return nil, fmt.Errorf("mpi: no device slice of %T", prototype)
	}
	var p unsafe.Pointer
	if rc := C.b200mpi_alloc(C.size_t(n*es), &p); rc != 0 {
		return nil, lastError(rc, 0)
	}
	return &DeviceSlice{Ptr: p, Len: n, dtype: dt}, nil
}

// Free returns the slice to the heap.
func (d *DeviceSlice) Free() { C.b200mpi_free(d.Ptr); d.Ptr = nil }

func lastError(rc C.int, tag int) error {
	msg := C.GoString(C.b200mpi_last_error())
	if rc == C.B200MPI_ERR_TAG_EXISTS {
		// the reference panics on a duplicate tag (network.go:469); surface it the same way
		panic(TagExists{Tag: tag}.Error())
	}
	return errors.New(msg)
}

func (c *Cuda) Init() error {
	if c.Password == "" {
		c.Password = FlagPassword
	}
	if c.Timeout == 0 {
		c.Timeout = time.Duration(FlagInitTimeout)
	}
	if c.Addr == "" {
		c.Addr = FlagAddr
	}
	if len(c.Addrs) == 0 {
		c.Addrs = append([]string(nil), FlagAllAddrs...)
	}
	gpu := FlagGpu
	if c.Gpu != nil {
		gpu = *c.Gpu
	}
	addr, all, pw := C.CString(c.Addr), C.CString(strings.Join(c.Addrs, ",")), C.CString(c.Password)
	defer C.free(unsafe.Pointer(addr))
	defer C.free(unsafe.Pointer(all))
	defer C.free(unsafe.Pointer(pw))
	if rc := C.b200mpi_init(addr, all, pw, C.int64_t(c.Timeout.Nanoseconds()), C.int(gpu)); rc != 0 {
		return lastError(rc, 0)
	}
	return nil
}

func (c *Cuda) Finalize() { C.b200mpi_finalize() }
func (c *Cuda) Rank() int { return int(C.b200mpi_rank()) }
func (c *Cuda) Size() int { return int(C.b200mpi_size()) }

// lower turns `data interface{}` into (pointer, count, dtype, memkind). Anything that is not one
// of the typed slices is gob-encoded to bytes first, which keeps strings and structs working.
func lower(data interface{}) (unsafe.Pointer, int, C.int, C.int, error) {
	switch v := data.(type) {
	case *DeviceSlice:
		return v.Ptr, v.Len, v.dtype, C.B200MPI_DEVICE, nil
	case []float64:
		return ptr(len(v), func() unsafe.Pointer { return unsafe.Pointer(&v[0]) }), len(v), dtF64, C.B200MPI_HOST, nil
	case []float32:
		return ptr(len(v), func() unsafe.Pointer { return unsafe.Pointer(&v[0]) }), len(v), dtF32, C.B200MPI_HOST, nil
	case []int64:
		return ptr(len(v), func() unsafe.Pointer { return unsafe.Pointer(&v[0]) }), len(v), dtI64, C.B200MPI_HOST, nil
	case []byte:
		return ptr(len(v), func() unsafe.Pointer { return unsafe.Pointer(&v[0]) }), len(v), dtU8, C.B200MPI_HOST, nil
	case Raw:
		return ptr(len(v), func() unsafe.Pointer { return unsafe.Pointer(&v[0]) }), len(v), dtU8, C.B200MPI_HOST, nil
	}
	var buf bytes.Buffer
	if err := gob.NewEncoder(&buf).Encode(data); err != nil {
		return nil, 0, 0, 0, err
	}
	b := buf.Bytes()
	return ptr(len(b), func() unsafe.Pointer { return unsafe.Pointer(&b[0]) }), len(b), dtU8, C.B200MPI_HOST, nil
}

func ptr(n int, f func() unsafe.Pointer) unsafe.Pointer {
	if n == 0 {
		return nil
	}
	return f()
}

func (c *Cuda) Send(data interface{}, destination, tag int) error {
	p, n, dt, kind, err := lower(data)
	if err != nil {
		return err
	}
	if rc := C.b200mpi_send(p, C.size_t(n), dt, C.int(destination), C.int(tag), kind); rc != 0 {
		return lastError(rc, tag)
	}
	return nil
}

// recvOnce is one b200mpi_recv call; n is the number of elements the message holds (also set
// when rc is B200MPI_ERR_TRUNCATE, in which case nothing was consumed and the call may be repeated
// with a larger buffer).
func recvOnce(p unsafe.Pointer, capacity int, dt, kind C.int, source, tag int) (n int, rc C.int) {
	var got C.size_t
	rc = C.b200mpi_recv(p, C.size_t(capacity), &got, dt, C.int(source), C.int(tag), kind)
	return int(got), rc
}

// recvBytes receives a byte message of unknown length into buf (grown when needed).
func recvBytes(buf []byte, source, tag int) ([]byte, error) {
	n, rc := recvOnce(ptr(len(buf), func() unsafe.Pointer { return unsafe.Pointer(&buf[0]) }), len(buf), dtU8, C.B200MPI_HOST, source, tag)
	if rc == C.B200MPI_ERR_TRUNCATE {
		buf = make([]byte, n)
		n, rc = recvOnce(unsafe.Pointer(&buf[0]), len(buf), dtU8, C.B200MPI_HOST, source, tag)
	}
	if rc != 0 {
		return nil, lastError(rc, tag)
	}
	return buf[:n], nil
}

// Receive takes a pointer to a slice (or a *DeviceSlice, or a pointer to any gob-decodable
// value). A slice that is too short is replaced by one of the sent length and the call repeated:
// the library keeps the message posted when it reports B200MPI_ERR_TRUNCATE. This is the
// resize-on-decode behaviour callers of the reference rely on (reference network.go:597).
func (c *Cuda) Receive(data interface{}, source, tag int) error {
	switch v := data.(type) {
	case *DeviceSlice:
		n, rc := recvOnce(v.Ptr, v.Len, v.dtype, C.B200MPI_DEVICE, source, tag)
		if rc != 0 {
			return lastError(rc, tag)
		}
		v.Len = n
		return nil
	case *[]float64:
		n, rc := recvOnce(ptr(len(*v), func() unsafe.Pointer { return unsafe.Pointer(&(*v)[0]) }), len(*v), dtF64, C.B200MPI_HOST, source, tag)
		if rc == C.B200MPI_ERR_TRUNCATE {
			*v = make([]float64, n)
			n, rc = recvOnce(unsafe.Pointer(&(*v)[0]), n, dtF64, C.B200MPI_HOST, source, tag)
		}
		if rc != 0 {
			return lastError(rc, tag)
		}
		*v = (*v)[:n]
		return nil
	case *[]float32:
		n, rc := recvOnce(ptr(len(*v), func() unsafe.Pointer { return unsafe.Pointer(&(*v)[0]) }), len(*v), dtF32, C.B200MPI_HOST, source, tag)
		if rc == C.B200MPI_ERR_TRUNCATE {
			*v = make([]float32, n)
			n, rc = recvOnce(unsafe.Pointer(&(*v)[0]), n, dtF32, C.B200MPI_HOST, source, tag)
		}
		if rc != 0 {
			return lastError(rc, tag)
		}
		*v = (*v)[:n]
		return nil
	case *[]int64:
		n, rc := recvOnce(ptr(len(*v), func() unsafe.Pointer { return unsafe.Pointer(&(*v)[0]) }), len(*v), dtI64, C.B200MPI_HOST, source, tag)
		if rc == C.B200MPI_ERR_TRUNCATE {
			*v = make([]int64, n)
			n, rc = recvOnce(unsafe.Pointer(&(*v)[0]), n, dtI64, C.B200MPI_HOST, source, tag)
		}
		if rc != 0 {
			return lastError(rc, tag)
		}
		*v = (*v)[:n]
		return nil
	case *[]byte:
		b, err := recvBytes(*v, source, tag)
		if err == nil {
			*v = b
		}
		return err
	case *Raw:
		b, err := recvBytes([]byte(*v), source, tag)
		if err == nil {
			*v = Raw(b)
		}
		return err
	}
	// anything else arrived gob-encoded (string, struct, ...): see lower()
	b, err := recvBytes(make([]byte, 256), source, tag)
	if err != nil {
		return err
	}
	return gob.NewDecoder(bytes.NewReader(b)).Decode(data)
}

// lowerTyped is lower() for the collectives: only typed slices, Raw and *DeviceSlice qualify. A
// gob-encoded temporary would receive the result instead of the caller's value (and its length
// may differ from rank to rank), so anything else is an error rather than a silent no-op.
func lowerTyped(what string, data interface{}) (unsafe.Pointer, int, C.int, C.int, error) {
	switch data.(type) {
	case *DeviceSlice, []float64, []float32, []int64, []byte, Raw:
		return lower(data)
	}
	return nil, 0, 0, 0, fmt.Errorf("mpi: %s needs []float64, []float32, []int64, []byte, Raw or *DeviceSlice, got %T", what, data)
}

func (c *Cuda) Bcast(data interface{}, root int) error {
	p, n, dt, kind, err := lowerTyped("Bcast", data)
	if err != nil {
		return err
	}
	if rc := C.b200mpi_bcast(p, C.size_t(n), dt, C.int(root), kind); rc != 0 {
		return lastError(rc, 0)
	}
	return nil
}

func (c *Cuda) Allreduce(send, recv interface{}, op Op) error {
	sp, sn, sdt, sk, err := lowerTyped("Allreduce", send)
	if err != nil {
		return err
	}
	rp, rn, rdt, rk, err := lowerTyped("Allreduce", recv)
	if err != nil {
		return err
	}
	if sn != rn || sdt != rdt || sk != rk {
		return errors.New("mpi: Allreduce send and recv must have the same type, length and memory kind")
	}
	if rc := C.b200mpi_allreduce(sp, rp, C.size_t(sn), sdt, C.int(op), sk); rc != 0 {
		return lastError(rc, 0)
	}
	return nil
}

func (c *Cuda) Allgather(send, recv interface{}) error {
	sp, sn, sdt, sk, err := lowerTyped("Allgather", send)
	if err != nil {
		return err
	}
	rp, rn, rdt, rk, err := lowerTyped("Allgather", recv)
	if err != nil {
		return err
	}
	if rn != sn*c.Size() || sdt != rdt || sk != rk {
		return errors.New("mpi: Allgather recv must hold Size()*len(send) elements of the same type")
	}
	if rc := C.b200mpi_allgather(sp, rp, C.size_t(sn), sdt, sk); rc != 0 {
		return lastError(rc, 0)
	}
	return nil
}

// ReduceScatter: send holds Size() blocks of len(recv); rank j receives the reduction of block j.
func (c *Cuda) ReduceScatter(send, recv interface{}, op Op) error {
	sp, sn, sdt, sk, err := lowerTyped("ReduceScatter", send)
	if err != nil {
		return err
	}
	rp, rn, rdt, rk, err := lowerTyped("ReduceScatter", recv)
	if err != nil {
		return err
	}
	if sn != rn*c.Size() || sdt != rdt || sk != rk {
		return errors.New("mpi: ReduceScatter send must hold Size()*len(recv) elements of the same type")
	}
	if rc := C.b200mpi_reduce_scatter(sp, rp, C.size_t(rn), sdt, C.int(op), sk); rc != 0 {
		return lastError(rc, 0)
	}
	return nil
}

// Reduce: Allreduce whose result lands on root only (recv may be nil elsewhere).
func (c *Cuda) Reduce(send, recv interface{}, op Op, root int) error {
	sp, sn, sdt, sk, err := lowerTyped("Reduce", send)
	if err != nil {
		return err
	}
	var rp unsafe.Pointer
	if recv != nil {
		var rn int
		var rdt, rk C.int
		rp, rn, rdt, rk, err = lowerTyped("Reduce", recv)
		if err != nil {
			return err
		}
		if sn != rn || sdt != rdt || sk != rk {
			return errors.New("mpi: Reduce send and recv must have the same type, length and memory kind")
		}
	}
	if rc := C.b200mpi_reduce(sp, rp, C.size_t(sn), sdt, C.int(op), C.int(root), sk); rc != 0 {
		return lastError(rc, 0)
	}
	return nil
}

// Alltoall: block j of send becomes block Rank() of rank j's recv.
func (c *Cuda) Alltoall(send, recv interface{}) error {
	sp, sn, sdt, sk, err := lowerTyped("Alltoall", send)
	if err != nil {
		return err
	}
	rp, rn, rdt, rk, err := lowerTyped("Alltoall", recv)
	if err != nil {
		return err
	}
	if sn != rn || sn%c.Size() != 0 || sdt != rdt || sk != rk {
		return errors.New("mpi: Alltoall send and recv must both hold Size() equal blocks of the same type")
	}
	if rc := C.b200mpi_alltoall(sp, rp, C.size_t(sn/c.Size()), sdt, sk); rc != 0 {
		return lastError(rc, 0)
	}
	return nil
}

// Isend is the Send of the design the reference sketches and comments out (mpi.go:132-143): it
// returns once data may be modified again, without waiting for the receiver.
func (c *Cuda) Isend(data interface{}, destination, tag int) error {
	p, n, dt, kind, err := lower(data)
	if err != nil {
		return err
	}
	if rc := C.b200mpi_isend(p, C.size_t(n), dt, C.int(destination), C.int(tag), kind); rc != 0 {
		return lastError(rc, tag)
	}
	return nil
}

// Wait blocks until destination confirmed the message sent with tag and frees the pair (mpi.go:146-152).
func (c *Cuda) Wait(destination, tag int) error {
	if rc := C.b200mpi_wait(C.int(destination), C.int(tag)); rc != 0 {
		return lastError(rc, tag)
	}
	return nil
}

func (c *Cuda) Barrier() error {
	if rc := C.b200mpi_barrier(); rc != 0 {
		return lastError(rc, 0)
	}
	return nil
}
