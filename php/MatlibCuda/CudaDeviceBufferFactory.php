<?php
namespace Rindow\Math\Matrix\Drivers\MatlibCuda;

/**
 * `$service->buffer(Service::LV_ACCELERATED)` (src/Drivers/AbstractMatlibService.php:316-318): the factory
 * NDArrayCL::newBuffer calls as Buffer($context, $bytes, $flags, $hostBuffer, $hostOffset, $dtype)
 * (src/NDArrayCL.php:147-168).
 */
class CudaDeviceBufferFactory
{
    public function isAvailable() : bool { return CudaFFI::isAvailable(); }

    public function Buffer(object $context, int $bytes, int $flags=0, ?object $hostBuffer=null, int $hostOffset=0,
        ?int $dtype=null) : object
    {
        return new CudaDeviceBuffer($context, $bytes, $flags, $hostBuffer, $hostOffset, $dtype);
    }
}
