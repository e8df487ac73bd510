<?php
namespace RindowTest\Math\Matrix\Drivers\MatlibCuda\CudaBlasTest;

use Rindow\Math\Matrix\MatrixOperator;
use Rindow\Math\Matrix\Drivers\Service;
use Rindow\Math\Matrix\Drivers\MatlibCuda\MatlibCuda;
// the reference's test classes are not autoloaded (cf. LinearAlgebraPHPModeTest.php:12-14)
if(!class_exists('RindowTest\Math\Matrix\Drivers\MatlibPHP\PhpBlasTest\PhpBlasTest')) {
    require_once __DIR__.'/../MatlibPHP/PhpBlasTest.php';
}
use RindowTest\Math\Matrix\Drivers\MatlibPHP\PhpBlasTest\PhpBlasTest as ORGTest;

/**
 * The reference's driver-level BLAS suite (PhpBlasTest.php, 4307 lines: raw (n, buf, off, ld) calls
 * and the buffer-overflow messages of :1988-2155) over CudaBlas -- same override point as
 * PhpBlasPHPModeTest.php:13-18.
 */
class CudaBlasTest extends ORGTest
{
    // the suite builds its arrays through $this->mo (trait Utils, tests/.../MatlibPHP/Utils.php:14-46)
    public function setUp() : void
    {
        $service = new MatlibCuda();
        if($service->serviceLevel()<Service::LV_ADVANCED) {
            $this->markTestSkipped('librindow_b200.so / B200 not available');
        }
        $this->mo = new MatrixOperator(service:$service);
    }

    public function getBlas()
    {
        return $this->mo->service()->blas();
    }
}
