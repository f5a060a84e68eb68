module kat
    implicit none
    real(kind=realType) :: acc
    integer(kind=intType) :: counter
    real(kind=realType), dimension(:, :, :), allocatable :: qq
contains

    subroutine scalars(n, x, res)
        ! integer / real arithmetic, integer powers, intrinsics, do with negative step, if / else if, select case
        integer(kind=intType), intent(in) :: n
        real(kind=realType), intent(in) :: x
        real(kind=realType), dimension(12), intent(out) :: res
        integer(kind=intType) :: i, k
        real(kind=realType) :: s
        res(1) = x**3 + x**(-2)
        res(2) = real(n / 3, realType) + mod(n, 5)
        res(3) = max(x, 2.5_realType, -x) - min(x, 0.25_realType)
        res(4) = sign(3.0_realType, -x) + abs(-x) + dim(x, 1.0_realType) + dim(1.0_realType, x)
        s = 0.0_realType
        do i = n, 1, -2
            s = s + real(i, realType) * 0.5_realType
        end do
        res(5) = s
        if (x > 2.0_realType) then
            res(6) = 1.0_realType
        else if (x > 1.0_realType .and. .not. (n == 3)) then
            res(6) = 2.0_realType
        else
            res(6) = 3.0_realType
        end if
        select case (n)
        case (1, 2)
            res(7) = 10.0_realType
        case (7)
            res(7) = 70.0_realType
        case default
            res(7) = -1.0_realType
        end select
        k = 0
        do i = 1, 10
            if (mod(i, 2) == 0) cycle
            if (i > 7) exit
            k = k + i
        end do
        res(8) = real(k, realType)
        res(9) = sqrt(x) * exp(-x) + log(x + 1.0_realType)
        call third_(s)
        res(10) = 1e-4 * x**2 + 5 * s
        res(11) = x**10
        res(12) = (x + 1.0_realType)**2 / 2 + 1 / (x + 1.0_realType)
    contains
        subroutine third_(t)
            real(kind=realType), intent(out) :: t
            t = 1.0_realType / 3.0_realType + x * 0.0_realType
        end subroutine third_
    end subroutine scalars

    subroutine arrays(n, a, out)
        ! lower bounds /= 1, column-major order, whole-array and section assignment, pointer sections with lower bound 1
        integer(kind=intType), intent(in) :: n
        real(kind=realType), dimension(0:n, -1:2), intent(inout) :: a
        real(kind=realType), dimension(8), intent(out) :: out
        real(kind=realType), dimension(0:n, -1:2), target :: b
        real(kind=realType), dimension(:, :), pointer :: p
        real(kind=realType), dimension(:), pointer :: q
        integer(kind=intType) :: i, j
        b = 1.5_realType
        do j = -1, 2
            do i = 0, n
                b(i, j) = b(i, j) + a(i, j) * real(j, realType)
            end do
        end do
        a(:, 0) = b(:, 1) * 2.0_realType
        out(1) = a(n, 0)
        out(2) = b(0, -1) + b(n, 2)
        p => b(1:, 0:)
        out(3) = p(1, 1)          ! = b(1, 0)
        out(4) = p(n, 3)          ! = b(n, 2)
        q => b(2, :)
        out(5) = q(1) + q(4)      ! = b(2,-1) + b(2,2)
        p => b
        out(6) = p(0, -1)         ! bounds of the target are kept
        call sum2(b(0, 1), b(1, 1), out(7))
        counter = counter + 1
        acc = acc + out(1)
        out(8) = acc + real(counter, realType)
    contains
        subroutine sum2(u, v, r)
            real(kind=realType), intent(in) :: u, v
            real(kind=realType), intent(out) :: r
            r = u + v + real(n, realType) * 0.0_realType
        end subroutine sum2
    end subroutine arrays

    subroutine optional_and_shape(m, n1, n2, res, flag)
        ! optional dummy + assumed-shape rank-2 dummy
        real(kind=realType), dimension(:, :), intent(inout) :: m
        integer(kind=intType), intent(in) :: n1, n2
        logical, intent(in), optional :: flag
        real(kind=realType), intent(out) :: res
        integer(kind=intType) :: i, j
        logical :: f
        f = .false.
        if (present(flag)) f = flag
        res = 0.0_realType
        do j = 1, n2
            do i = 1, n1
                if (f) m(i, j) = -m(i, j)
                res = res + m(i, j) * real(i + 10 * j, realType)
            end do
        end do
    end subroutine optional_and_shape

    subroutine driver_shape(res)
        real(kind=realType), dimension(2), intent(out) :: res
        real(kind=realType), dimension(3, 2) :: m
        integer(kind=intType) :: i, j
        do j = 1, 2
            do i = 1, 3
                m(i, j) = real(i, realType) + 0.1_realType * real(j, realType)
            end do
        end do
        call optional_and_shape(m, 3, 2, res(1))
        call optional_and_shape(m, 3, 2, res(2), .true.)
    end subroutine driver_shape
    subroutine alloc_case(n, res)
        ! module-level allocatable with explicit lower bounds (sa_block: allocate(qq(2:il, 2:jl, 2:kl))), named constructs,
        ! integer min/max
        integer(kind=intType), intent(in) :: n
        real(kind=realType), dimension(4), intent(out) :: res
        integer(kind=intType) :: i, j, k, m
        allocate (qq(2:n, 2:n + 1, 0:1))
        do k = 0, 1
            do j = 2, n + 1
                do i = 2, n
                    qq(i, j, k) = real(100 * k + 10 * j + i, realType)
                end do
            end do
        end do
        res(1) = qq(2, 2, 0) + qq(n, n + 1, 1)
        m = 0
        outer: do j = 2, n + 1
            testinner: if (j > 3) then
                exit
            end if testinner
            m = m + j
        end do outer
        res(2) = real(m, realType)
        res(3) = real(max(n, 3) - min(n, 3) + mod(-7, 3), realType)
        res(4) = qq(n, 2, 1) * 0.5_realType
        deallocate (qq)
    end subroutine alloc_case
#ifdef NEVER
    subroutine broken(
#endif
end module kat
