<?php
namespace RindowTest\Math\Matrix\NDArrayCudaTest;

use PHPUnit\Framework\TestCase;
use Rindow\Math\Matrix\NDArrayPhp;
use Rindow\Math\Matrix\NDArrayCuda;
use Rindow\Math\Matrix\Drivers\Service;
use Rindow\Math\Matrix\Drivers\MatlibCuda\MatlibCuda;
use Interop\Polite\Math\Matrix\NDArray;
use Interop\Polite\Math\Matrix\OpenCL;
use Interop\Polite\Math\Matrix\DeviceBuffer;
use function Rindow\Math\Matrix\R;

/**
 * tests/RindowTest/Math/Matrix/NDArrayCLTest.php (:56-160) over the B200 driver: the same assertions with
 * NDArrayCuda / CudaQueue in place of NDArrayCL / the OpenCL queue.  The executable transcription of this file is
 * tests/test_gpu_device_ndarray.py of the B200 repository.
 */
class NDArrayCudaTest extends TestCase
{
    protected $service;

    public function setUp() : void
    {
        $this->service = new MatlibCuda();
        if($this->service->serviceLevel()<Service::LV_ACCELERATED) {
            $this->markTestSkipped('librindow_b200.so / B200 not available');
        }
    }

    protected function device(NDArray $host, int $flags)
    {
        $queue = $this->service->createQueue();
        return [$queue, new NDArrayCuda($queue, $host->buffer(), $host->dtype(), $host->shape(), $host->offset(),
            $flags|OpenCL::CL_MEM_COPY_HOST_PTR, service:$this->service)];
    }

    public function testSimpleArrayoffsetGet()
    {
        $host = new NDArrayPhp([[1,2],[3,4],[5,6]], service:$this->service);
        [$queue, $array] = $this->device($host, OpenCL::CL_MEM_READ_ONLY);
        $this->assertEquals([3,2], $array->shape());
        $this->assertEquals(NDArray::float32, $array->dtype());
        $this->assertEquals(6, $array->size());
        $this->assertInstanceof(DeviceBuffer::class, $array->buffer());
        $this->assertEquals(6*32/8, $array->buffer()->bytes());
        $this->assertEquals([3,4], $array[1]->toNDArray()->toArray());
        $this->assertEquals([[3,4],[5,6]], $array[R(1,3)]->toNDArray()->toArray());
        $this->assertEquals(0, $array[1][1]->ndim());
        $this->assertEquals(4, $array[1][1]->toNDArray()->toArray());
    }

    public function testSimpleArrayoffsetSet()
    {
        $host = new NDArrayPhp([[1,2],[3,4],[5,6]], service:$this->service);
        [$queue, $array] = $this->device($host, OpenCL::CL_MEM_READ_WRITE);
        [, $src] = $this->device(new NDArrayPhp([11,12], service:$this->service), OpenCL::CL_MEM_READ_ONLY);
        $array[1] = $src;
        $queue->finish();
        $this->assertEquals([[1,2],[11,12],[5,6]], $array->toNDArray()->toArray());
        [, $s] = $this->device(new NDArrayPhp(4, service:$this->service), OpenCL::CL_MEM_READ_ONLY);
        $array[1][1] = $s;
        $queue->finish();
        $this->assertEquals([[1,2],[11,4],[5,6]], $array->toNDArray()->toArray());
    }

    public function testClone()
    {
        $host = new NDArrayPhp([[1,2],[3,4],[5,6]], dtype:NDArray::int32, service:$this->service);
        [$queue, $array] = $this->device($host, OpenCL::CL_MEM_READ_ONLY);
        $pattern = new NDArrayPhp([[0,0],[0,0],[0,0]], dtype:NDArray::int32, service:$this->service);
        $array2 = clone $array;
        $array->buffer()->write($queue, $pattern->buffer());
        $this->assertEquals([[0,0],[0,0],[0,0]], $array->toArray());
        $this->assertEquals([[1,2],[3,4],[5,6]], $array2->toArray());
    }
}
