// worker/gpu: the GPU worker of MWindels/distributed-raytracer = worker/distributed/main.go with the per-pixel loop
// (lines 67-88) replaced by one call into the B200 backend.  Registration, heartbeats, the 2 s watchdog and the gRPC
// service are the reference's own, unchanged; the master (master/pool, master/registrar.go) cannot tell the difference.
// NOT COMPILED in the repository this file ships in (no Go toolchain there).
//
//	usage: gpu <master address> <listening port> [cuda device] [exact]
package main

import (
	"bytes"
	"context"
	"encoding/gob"
	"fmt"
	"log"
	"net"
	"os"
	"strconv"
	"time"

	"github.com/golang/protobuf/ptypes/empty"
	"github.com/mwindels/distributed-raytracer/shared/comms"
	"github.com/mwindels/distributed-raytracer/shared/state"
	"github.com/mwindels/distributed-raytracer/worker/gpu/rtb200"
	"google.golang.org/grpc"
	"google.golang.org/grpc/codes"
	"google.golang.org/grpc/status"
)

const registerFrequency uint = 500 // worker/distributed/main.go:21
const traceTimeout uint = 2000     // worker/distributed/main.go:24

// Tracer implements the comms.TraceServer interface (worker/distributed/main.go:27-32) on one GPU.
type Tracer struct {
	scene                     state.Environment
	gpu                       *rtb200.Scene // the registered scene on the device: geometry + BVH + lanes
	screenWidth, screenHeight uint
	resetTraceTimeout         chan struct{}
}

func (t *Tracer) timeoutReset() {
	defer func() { recover() }()
	t.resetTraceTimeout <- struct{}{}
}

// BulkTrace traces a batch of rays (worker/distributed/main.go:46-91).
func (t *Tracer) BulkTrace(ctx context.Context, req *comms.WorkOrder) (*comms.TraceResults, error) {
	t.timeoutReset()
	width, height := int(req.GetWidth()), int(req.GetHeight())
	results := &comms.TraceResults{Results: make([]*comms.TraceResults_Colour, width*height, width*height)}

	// Decode the mutable state for this frame (:56-64).
	var diff state.EnvMutables
	if req.GetDiff() != nil {
		if err := gob.NewDecoder(bytes.NewBuffer(req.GetDiff())).Decode(&diff); err != nil {
			return nil, err
		}
		diff.LinkTo(t.scene)
	}
	if err := ctx.Err(); err == context.Canceled { // :73, once per order instead of once per pixel
		return nil, err
	}

	// The loop of :67-88 as one call: buf[(i*height+j)*3 ...] = RGB() of pixel (x+i, y+j), misses black.
	buf := make([]byte, width*height*3)
	if err := t.gpu.BulkTrace(&diff, uint32(t.screenWidth), uint32(t.screenHeight),
		req.GetX(), req.GetY(), req.GetWidth(), req.GetHeight(), buf); err != nil {
		return nil, status.Error(codes.Internal, err.Error())
	}
	colours := make([]comms.TraceResults_Colour, width*height) // one allocation instead of width*height
	for k := range results.Results {
		colours[k].R, colours[k].G, colours[k].B = uint32(buf[3*k]), uint32(buf[3*k+1]), uint32(buf[3*k+2])
		results.Results[k] = &colours[k]
	}
	return results, nil
}

// Heartbeat keeps the worker from disconnecting from the master (worker/distributed/main.go:94-98).
func (t *Tracer) Heartbeat(ctx context.Context, req *empty.Empty) (*empty.Empty, error) {
	t.timeoutReset()
	return &empty.Empty{}, nil
}

// register registers this worker with the master and uploads the scene to the GPU (worker/distributed/main.go:101-129).
func register(registerAddr string, listenPort uint32, device int, fast bool) (*Tracer, error) {
	conn, err := grpc.Dial(registerAddr, grpc.WithInsecure())
	if err != nil {
		return nil, err
	}
	defer conn.Close()
	client := comms.NewRegistrationClient(conn)
	stateMsg, err := client.Register(context.Background(), &comms.WorkerLink{Port: listenPort})
	if err != nil {
		return nil, err
	}
	var newScene state.Environment
	if stateMsg.GetState() == nil {
		return nil, fmt.Errorf("No scene data recieved.")
	}
	if err = gob.NewDecoder(bytes.NewBuffer(stateMsg.GetState())).Decode(&newScene); err != nil {
		return nil, err
	}
	gpu, err := rtb200.NewScene(newScene, device, fast) // immutable meshes: upload + on-GPU BVH build, once
	if err != nil {
		return nil, err
	}
	return &Tracer{scene: newScene, gpu: gpu, screenWidth: uint(stateMsg.GetScreenWidth()),
		screenHeight: uint(stateMsg.GetScreenHeight()), resetTraceTimeout: make(chan struct{})}, nil
}

func main() {
	if len(os.Args) < 3 {
		log.Fatalln("Improper parameters.  This program requires the parameters:" +
			"\n\t(1) master address (including port)" +
			"\n\t(2) work order listening port" +
			"\n\t(3) CUDA device ordinal (optional, default 0)" +
			"\n\t(4) \"exact\" for the float64 bit-exact mode (optional, default fast)")
	}
	masterAddr := os.Args[1]
	orderPort, err := strconv.ParseUint(os.Args[2], 10, 32)
	if err != nil {
		log.Fatalf("Could not parse port number \"%s\": %v.\n", os.Args[2], err)
	}
	device, fast := 0, true
	if len(os.Args) > 3 {
		if device, err = strconv.Atoi(os.Args[3]); err != nil {
			log.Fatalf("Could not parse device ordinal \"%s\": %v.\n", os.Args[3], err)
		}
	}
	if len(os.Args) > 4 && os.Args[4] == "exact" {
		fast = false
	}

	for {
		tracer, err := register(masterAddr, uint32(orderPort), device, fast)
		if err == nil {
			server := grpc.NewServer()
			comms.RegisterTraceServer(server, tracer)
			listener, err := net.Listen("tcp", fmt.Sprintf(":%d", orderPort))
			if err != nil {
				log.Fatalf("Failed to listen on port \"%d\": %v.\n", orderPort, err)
			}
			go func() { // the watchdog of worker/distributed/main.go:161-172
				for {
					select {
					case <-tracer.resetTraceTimeout:
					case <-time.After(time.Millisecond * time.Duration(traceTimeout)):
						close(tracer.resetTraceTimeout)
						server.GracefulStop()
						return
					}
				}
			}()
			if err = server.Serve(listener); err != nil {
				log.Printf("Tracer interrupted: %v.\n", err)
			} else {
				log.Printf("Tracer timed out after recieving no orders or heartbeats.\n")
			}
			tracer.gpu.Close() // re-registration = destroy + create
		} else {
			log.Printf("Failed to register: %v.\n", err)
		}
		time.Sleep(time.Millisecond * time.Duration(registerFrequency))
	}
}
