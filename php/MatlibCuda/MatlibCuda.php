<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

use Rindow\Math\Matrix\Drivers\AbstractMatlibService;
use Rindow\Math\Matrix\Drivers\Service;
use LogicException;

/**
 * The B200 service: an AbstractMatlibService with the three CUDA factories injected
 * (src/Drivers/AbstractMatlibService.php:43-75).  The three factories make the level LV_ADVANCED, at which
 * NDArrayPhp / LinearAlgebra / MatrixOperator use the CUDA driver with zero edits:
 *
 *     $mo = new MatrixOperator(service: new MatlibCuda());
 *     // or: new MatrixOperator(catalog: [MatlibCuda::class])      (src/Drivers/Selector.php:36-71)
 *
 * and, as the same library also provides device-resident buffers and a queue, the accelerated tier (:193-218, there:
 * OpenCL + CLBlast factories) is diagnosed too: buffer(LV_ACCELERATED) is the CudaDeviceBuffer factory NDArrayCuda
 * asks for, createQueue() the CudaQueue NDArrayCuda / LinearAlgebraCuda are constructed with:
 *
 *     $la = (new MatrixOperatorCuda(service: new MatlibCuda()))->laAccelerated('cuda');
 */
class MatlibCuda extends AbstractMatlibService
{
    protected string $name = 'matlib_cuda';

    protected function injectDefaultFactories() : void
    {
        $this->bufferFactory   ??= new CudaBufferFactory();
        $this->openblasFactory ??= new CudaBlasFactory();
        $this->mathFactory     ??= new CudaMathFactory();
    }

    protected ?CudaDeviceBufferFactory $deviceBufferFactory = null;

    public function serviceLevel() : int
    {
        $level = parent::serviceLevel();
        if($level==Service::LV_ADVANCED && ($this->bufferFactory instanceof CudaBufferFactory)) {
            $this->serviceLevel = $level = Service::LV_ACCELERATED;
        }
        return $level;
    }

    public function buffer(?int $level=null) : object
    {
        if($level==Service::LV_ACCELERATED) {
            return $this->deviceBufferFactory ??= new CudaDeviceBufferFactory();
        }
        return parent::buffer($level);
    }

    /** @param array<string,mixed> $options */
    public function createQueue(?array $options=null) : object
    {
        if($this->serviceLevel()<Service::LV_ACCELERATED) {
            throw new LogicException('librindow_b200.so / B200 not available');
        }
        return new CudaQueue($options);
    }

    /** the device-tier driver objects are the same CudaBlas / CudaMath: they take device buffers positionally */
    public function blasCuda(object $queue) : object { return $this->blas(); }
    public function mathCuda(object $queue) : object { return $this->math(); }

    public function info() : string
    {
        // AbstractMatlibService::info() prints the OpenCL factories at LV_ACCELERATED (:241-244)
        $info = "Service Level   : Accelerated\n";
        $info .= "Buffer Factory  : ".get_class($this->buffer())."\n";
        $info .= "BLAS Driver     : ".get_class($this->blas())."\n";
        $info .= "Math Driver     : ".get_class($this->math())."\n";
        $info .= "Device Buffers  : ".CudaDeviceBufferFactory::class."\n";
        return $info;
    }
}
