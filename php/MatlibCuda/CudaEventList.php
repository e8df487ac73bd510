<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use FFI;
use Countable;

/**
 * Stream-ordered completion markers: the stand-in for the OpenCL EventList the reference threads through its device
 * calls as ($events, $waitEvents) (src/Drivers/MatlibCL/OpenCLMath.php:1851-1859, src/LinearAlgebraCL.php:1867-1875).
 * record() marks "everything queued on the library stream so far" (rb200_event_record); wait() blocks the host.
 */
class CudaEventList implements Countable
{
    /** @var array<int,mixed> */
    protected array $events = [];

    public function record() : void
    {
        $ffi = CudaFFI::lib();
        $ev = $ffi->new('void*');
        CudaFFI::check($ffi->rb200_event_create(FFI::addr($ev)));
        CudaFFI::check($ffi->rb200_event_record($ev));
        $this->events[] = $ev;
    }

    public function copy(CudaEventList $other) : void
    {
        foreach($other->events as $ev) { $this->events[] = $ev; }
    }

    public function count() : int { return count($this->events); }

    public function wait() : void
    {
        $ffi = CudaFFI::lib();
        foreach($this->events as $ev) { CudaFFI::check($ffi->rb200_event_synchronize($ev)); }
    }

    public function done() : bool
    {
        $ffi = CudaFFI::lib();
        $flag = $ffi->new('int');
        foreach($this->events as $ev) {
            CudaFFI::check($ffi->rb200_event_query($ev, FFI::addr($flag)));
            if(!$flag->cdata) { return false; }
        }
        return true;
    }
}
