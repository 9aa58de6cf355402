// Optional (SURVEY s8 f2): answer BulkTrace with the reply already in protobuf wire form, encoded on the GPU, instead of
// building width*height comms.TraceResults_Colour messages and letting grpc-go marshal them.  The master is untouched: it
// receives an ordinary comms.TraceResults.  NOT COMPILED in the repository this file ships in.
//
// Use: the generated comms.TraceServer interface fixes BulkTrace's return type to *comms.TraceResults, so the service is
// registered with a hand-written grpc.ServiceDesc (a copy of the generated _Trace_serviceDesc) whose BulkTrace handler
// returns &RawTraceResults{Wire: wire} with wire from rtb200.Scene.BulkTraceEncoded; the codec below passes it through.
package main

import (
	"google.golang.org/grpc/encoding"
	_ "google.golang.org/grpc/encoding/proto" // make sure the stock codec is registered first
)

// RawTraceResults carries the wire bytes of a comms.TraceResults message (rtb200.Scene.BulkTraceEncoded).
type RawTraceResults struct{ Wire []byte }

type passThroughCodec struct{ encoding.Codec }

// Marshal hands pre-encoded replies to the transport as they are and everything else to the stock "proto" codec.
func (c passThroughCodec) Marshal(v interface{}) ([]byte, error) {
	if r, ok := v.(*RawTraceResults); ok {
		return r.Wire, nil
	}
	return c.Codec.Marshal(v)
}

func init() {
	// same Name() ("proto") as the codec it wraps: replaces it for this process
	encoding.RegisterCodec(passThroughCodec{encoding.GetCodec("proto")})
}
