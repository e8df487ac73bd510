<?php
namespace Rindow\Math\Matrix;

use Countable;
use IteratorAggregate;
use Traversable;
use InvalidArgumentException;
use OutOfRangeException;
use LogicException;
use Interop\Polite\Math\Matrix\NDArray;
use Interop\Polite\Math\Matrix\Buffer;
use Interop\Polite\Math\Matrix\LinearBuffer;
use Interop\Polite\Math\Matrix\DeviceBuffer;
use Interop\Polite\Math\Matrix\OpenCL;
use Rindow\Math\Matrix\Drivers\Service;

/**
 * Device-resident N-d array of the B200 driver, "alongside NDArrayCL": same constructor, same view / copy / read-back
 * rules (src/NDArrayCL.php:68-140 ctor, :252-272 toNDArray, :274-388 offsetGet, :390-442 offsetSet, :444-515
 * unsupported operations and clone), over a CudaDeviceBuffer.  Drop into src/ next to NDArrayCL.php.
 * Executable mirror with tests: rindow_math_matrix_b200/ndarray_cuda.py, tests/test_gpu_device_ndarray.py.
 *
 * @implements IteratorAggregate<int, mixed>
 */
class NDArrayCuda extends NDArrayCL
{
    // NDArrayCL's constructor is final and asks `$queue->getContext()` and `$service->buffer(LV_ACCELERATED)` for the
    // context and the buffer factory: CudaQueue and MatlibCuda answer both, so the inherited constructor, shape(),
    // reshape(), offsetGet()/offsetSet() (buffer->copy with byte offsets), toNDArray() (buffer->read) and getIterator()
    // work unchanged.  Only clone differs: it must not go through `$service->openCL()->EventList()`.

    public function __clone()
    {
        if(!($this->buffer instanceof DeviceBuffer)) {
            throw new \RuntimeException('Unknown buffer type is uncloneable:'.get_class($this->buffer));
        }
        $flags = $this->flags & ~OpenCL::CL_MEM_COPY_HOST_PTR;
        $newBuffer = $this->clBufferFactory->Buffer(
            $this->context, $this->buffer->bytes(), $flags, null, 0, $this->buffer->dtype());
        $newBuffer->copy($this->queue, $this->buffer, 0, 0, 0);
        $this->queue->finish();
        $this->flags = $flags;
        $this->buffer = $newBuffer;
    }
}
